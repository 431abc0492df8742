// divsqrt_probe.cu -- test kernels for tests/test_gpu_divsqrt.py (not product): the check-free division / square root of
// rn_prelude.cuh (rn_div_nc, rn_sqrt_nc) against CUDA's own operators, and the straight-line common paths of the fdlibm
// functions (rn_strict_exp / log / pow: speculative form, coefficients from the constant bank) against the complete
// transcriptions (rn_strict_*_full), all ON THE DEVICE and bit for bit.  Built with --fmad=false like the parity kernels.
#include "../../rainier_b200/csrc/rn_prelude.cuh"

__device__ __forceinline__ bool same_bits(double a, double b) {
  return __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b);
}
// which: 0 a/b, 1 sqrt(a), 2 exp(a), 3 log(a), 4 pow(a, b), 5 pow(a, -2.0) (literal exponent: the y tests fold)
extern "C" __global__ void divsqrt_probe_kernel(int which, const double* a, const double* b, long long n, unsigned long long* mismatches,
                                                long long* first) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = a[i], y = b ? b[i] : 0.0;
    double got, want;
    switch (which) {
      case 0: got = rn_div_nc(x, y); want = x / y; break;
      case 1: got = rn_sqrt_nc(x); want = sqrt(x); break;
      case 2: got = rn_strict_exp(x); want = rn_strict_exp_full(x); break;
      case 3: got = rn_strict_log(x); want = rn_strict_log_full(x); break;
      case 4: got = rn_strict_pow(x, y); want = rn_strict_pow_full(x, y); break;
      default: got = rn_strict_pow(x, -2.0); want = rn_strict_pow_full(x, -2.0); break;
    }
    if (!same_bits(got, want)) {
      bad += 1;
      atomicMin((unsigned long long*)first, (unsigned long long)i);
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}
// row functions (tolerance class): which 0 rn_row_exp vs exp, 1 rn_row_log vs log, 2 rn_row_rcp vs 1.0 / x.  Results must be the
// same class (NaN / +-inf / +-0 identical); finite results are compared in units in the last place: max_ulp receives the largest
// distance, and results that differ by more than `tol_ulp` are counted.
extern "C" __global__ void rowlibm_probe_kernel(int which, const double* a, long long n, long long tol_ulp, unsigned long long* mismatches,
                                                unsigned long long* max_ulp, long long* first) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long bad = 0, worst = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = a[i];
    double got, want;
    switch (which) {
      case 0: got = rn_row_exp(x); want = exp(x); break;
      case 1: got = rn_row_log(x); want = log(x); break;
      default: got = rn_row_rcp(x); want = 1.0 / x; break;
    }
    bool ok;
    if (want != want || got != got) ok = (want != want) && (got != got);
    else if (isinf(want) || isinf(got) || want == 0.0 || got == 0.0) {
      ok = __double_as_longlong(got) == __double_as_longlong(want);
      // subnormal neighbourhood of zero (exp underflow, 1/huge): one unit of the subnormal grid is allowed
      if (!ok && fabs(want) < 2.3e-308 && fabs(got) < 2.3e-308) ok = llabs(__double_as_longlong(got) - __double_as_longlong(want)) <= tol_ulp;
    } else {
      const long long d = llabs(__double_as_longlong(got) - __double_as_longlong(want));  // same sign is implied by d small
      ok = d <= tol_ulp;
      if ((unsigned long long)d > worst && d < (1LL << 40)) worst = (unsigned long long)d;
    }
    if (!ok) {
      bad += 1;
      atomicMin((unsigned long long*)first, (unsigned long long)i);
    }
  }
  if (bad) atomicAdd(mismatches, bad);
  atomicMax(max_ulp, worst);
}
extern "C" int rowlibm_probe_run(int which, const double* a, long long n, long long tol_ulp, unsigned long long* mismatches,
                                 unsigned long long* max_ulp, long long* first) {
  rowlibm_probe_kernel<<<148 * 8, 256>>>(which, a, n, tol_ulp, mismatches, max_ulp, first);
  return (int)cudaDeviceSynchronize();
}
extern "C" int divsqrt_probe_run(int which, const double* a, const double* b, long long n, unsigned long long* mismatches, long long* first) {
  divsqrt_probe_kernel<<<148 * 8, 256>>>(which, a, b, n, mismatches, first);
  return (int)cudaDeviceSynchronize();
}
