// rn_diag.cuh -- on-device convergence diagnostics over a device-resident sample block: the reductions behind
// Trace.diagnostics (rainier-core/src/main/scala/com/stripe/rainier/core/Trace.scala:11-21,49-121: rHat / v of the
// Stan manual 30.3, variogram-based autocorrelation / effective sample size of 30.4).  With thousands of chains the
// reference's host-side loops (O(chains * iterations * lags) over List[List[Array[Double]]]) and the device->host copy
// of every sample dominate the call; here only [n][2] numbers leave the device.
//
//   rn_k_diag_chain   one thread per (parameter, chain): mean, variance and the variograms of lags 1..max_lag of that
//                     chain's series (summed over iterations in the reference's sequential order), reduced over the 128
//                     chains of the block in a fixed order.
//   rn_k_diag_reduce  fixed-shape tree sums over blocks / chains (deterministic), optionally of squared deviations.
// The scalar epilogue (b, w, v, rHat, the lag loop with its termination rule, ess) runs on the host in rn_runtime.cpp.
#ifndef RN_DIAG_CUH
#define RN_DIAG_CUH
#ifndef RN_HOST_EMULATION

// sample(t, i, c) = s[t*st + i*si + c*sc].  grid = (ceil(C/128), n), 128 threads: thread = chain, blockIdx.y = parameter.
// The chain's series is staged once in shared memory (x[t][thread], conflict-free) when iterations*128 doubles fit
// (`use_smem`), so the O(iterations * lags) variogram loops never touch global memory again.  Per block the 128 chains'
// contributions are added in a fixed order (warp shuffles, then 4 warp totals) and written as one partial per
// (quantity, parameter, block): partial[(q * n + i) * nblk + blockIdx.x], q = 0 mean, 1 variance, 2.. variogram(lag).
extern __shared__ double rn_diag_smem[];
RN_DEVICE double rn_diag_block_sum(double v, double* red4) {
  RN_UNROLL
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red4[threadIdx.x >> 5] = v;
  __syncthreads();
  return ((red4[0] + red4[1]) + red4[2]) + red4[3];
}
RN_GLOBAL void rn_k_diag_chain(const double* RN_RESTRICT s, long long st, long long si, long long sc, int I, int n, int C,
                               int max_lag, int use_smem, double* RN_RESTRICT mean, double* RN_RESTRICT partial) {
  __shared__ double red4[4];
  const int i = blockIdx.y, c = blockIdx.x * 128 + threadIdx.x, nblk = gridDim.x;
  const bool live = c < C;
  const double* xg = s + (long long)i * si + (long long)(live ? c : 0) * sc;
  double* xs = rn_diag_smem + threadIdx.x;  // x[t] at xs[t * 128]
  if (use_smem)
    for (int t = 0; t < I; t++) xs[(size_t)t * 128] = live ? xg[(long long)t * st] : 0.0;
#define RN_DIAG_X(t) (use_smem ? xs[(size_t)(t) * 128] : (live ? xg[(long long)(t) * st] : 0.0))
  double sum = 0.0;  // t.sum / n, Trace.scala:69-71 (sequential over iterations, like the reference)
  for (int t = 0; t < I; t++) sum += RN_DIAG_X(t);
  const double m = sum / (double)I;
  double ss = 0.0;  // t.map(a => pow(a - m, 2)).sum / (n - 1), Trace.scala:79-86
  for (int t = 0; t < I; t++) {
    const double d = RN_DIAG_X(t) - m;
    ss += d * d;
  }
  if (live) mean[(size_t)i * C + c] = m;
  double tot = rn_diag_block_sum(live ? m : 0.0, red4);
  if (threadIdx.x == 0) partial[((size_t)0 * n + i) * nblk + blockIdx.x] = tot;
  tot = rn_diag_block_sum(live ? ss / (double)(I - 1) : 0.0, red4);
  if (threadIdx.x == 0) partial[((size_t)1 * n + i) * nblk + blockIdx.x] = tot;
  for (int lag = 1; lag <= max_lag; lag++) {  // Trace.variogram, Trace.scala:111-119
    double v = 0.0;
    for (int t = lag; t < I; t++) {
      const double d = RN_DIAG_X(t) - RN_DIAG_X(t - lag);
      v += d * d;
    }
    tot = rn_diag_block_sum(live ? v / (double)(I - lag) : 0.0, red4);
    if (threadIdx.x == 0) partial[((size_t)(1 + lag) * n + i) * nblk + blockIdx.x] = tot;
  }
#undef RN_DIAG_X
}

// out[q] = sum_c f(in[q*C + c]),  f(x) = x or (x - shift[q])^2 ; one 256-thread block per q
RN_GLOBAL void rn_k_diag_reduce(const double* RN_RESTRICT in, int C, const double* RN_RESTRICT shift, double* RN_RESTRICT out) {
  __shared__ double red[256];
  const int q = blockIdx.x;
  const double sh = shift ? shift[q] : 0.0;
  double acc = 0.0;
  for (int c = threadIdx.x; c < C; c += 256) {
    const double x = in[(size_t)q * C + c];
    acc += shift ? (x - sh) * (x - sh) : x;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[q] = red[0];
}

#endif
#endif  // RN_DIAG_CUH
