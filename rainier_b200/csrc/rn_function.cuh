// rn_function.cuh -- hand-written kernel around the emitted rn_function(): batched evaluation of a compiled function.
//
// Replaces the per-sample loop of Generator.prepare (rainier-core/.../core/Generator.scala:76-93): for every posterior
// draw the reference calls CompiledFunction.output(cf, array, globalBuf, i) once per requirement i
// (rainier-compute/.../ir/CompiledFunction.scala:122-140), re-walking the generated methods each time.  Here one thread
// owns one draw ("point"), evaluates the straight-line SSA of ALL m outputs once (shared sub-expressions once, like the
// reference's globals) and stores them.  HBM-bound by construction: (RN_N + RN_M) * 8 algorithmic bytes per point.
//
// Addressing is generic so that the kernel reads posterior draws where they already are:
//   element i of point p:  x[(p / in_inner) * in_outer + (p % in_inner) * in_pstride + i * in_estride]
//     sampler layout [iteration][n][chain] (rn_sampler_run): in_inner = chains, in_outer = n*chains, in_pstride = 1,
//       in_estride = chains   -> consecutive threads read consecutive doubles (coalesced)
//     row-major [count][n] (host callers): in_inner = count, in_pstride = n, in_estride = 1
//   output j of point p:   out[(p / out_inner) * out_outer + (p % out_inner) * out_pstride + j * out_estride]
#ifndef RN_FUNCTION_CUH
#define RN_FUNCTION_CUH

// RnEvalArgs: rn_args.h

RN_GLOBAL void rn_k_eval(const RnEvalArgs A) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  int err = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < A.count; p += stride) {
    double q[RN_NQ];
    const double* RN_RESTRICT xp = A.x + (p / A.in_inner) * A.in_outer + (p % A.in_inner) * A.in_pstride;
    RN_UNROLL
    for (int i = 0; i < RN_N; i++) q[i] = RN_LDG(xp + (long long)i * A.in_estride);
    rn_function(q, A.out + (p / A.out_inner) * A.out_outer + (p % A.out_inner) * A.out_pstride, A.out_estride, err);
  }
#ifdef RN_HOST_EMULATION
  if (err) *A.err |= err;
#else
  if (err) atomicOr(A.err, err);
#endif
}

// Fixed-order row reduction of the values rn_k_eval wrote as [output][row]: sums[j] += sum over rows of vals[j][.] -- one block per
// output, thread t adds rows t, t + 256, ... in order, then a fixed tree (no atomics: the same bits every run).  Used by the
// device-side inlining of separable likelihoods (rn_inline.hpp), where the "outputs" are column-only monomials of a target.
#ifndef RN_HOST_EMULATION
RN_GLOBAL void rn_k_reduce_rows(const double* RN_RESTRICT vals, const long long rows, const int m, double* sums) {
  __shared__ double red[256];
  const int j = (int)blockIdx.x;
  if (j >= m) return;
  const double* v = vals + (size_t)j * (size_t)rows;
  double acc = 0.0;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) acc += v[r];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = (int)blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[j] += red[0];
}
#endif

#endif  // RN_FUNCTION_CUH
