// rn_diag.cuh -- on-device convergence diagnostics over a device-resident sample block: the reductions behind
// Trace.diagnostics (rainier-core/src/main/scala/com/stripe/rainier/core/Trace.scala:11-21,49-121: rHat / v of the
// Stan manual 30.3, variogram-based autocorrelation / effective sample size of 30.4).  With thousands of chains the
// reference's host-side loops (O(chains * iterations * lags) over List[List[Array[Double]]]) and the device->host copy
// of every sample dominate the call; here only [n][2] numbers leave the device.
//
//   rn_k_diag_chain   one thread per (parameter, chain): mean, variance and the variograms of lags 1..max_lag of that
//                     chain's series, summed over iterations in the reference's sequential order.
//   rn_k_diag_reduce  fixed-shape tree sums over chains (deterministic), optionally of squared deviations.
// The scalar epilogue (b, w, v, rHat, the lag loop with its termination rule, ess) runs on the host in rn_runtime.cpp.
#ifndef RN_DIAG_CUH
#define RN_DIAG_CUH
#ifndef RN_HOST_EMULATION

// sample(t, i, c) = s[t*st + i*si + c*sc];  outputs are [.][n][C], chain fastest
RN_GLOBAL void rn_k_diag_chain(const double* RN_RESTRICT s, long long st, long long si, long long sc, int I, int n, int C,
                               int max_lag, int i_fastest, double* RN_RESTRICT mean, double* RN_RESTRICT var,
                               double* RN_RESTRICT vario) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * C) return;
  const int i = i_fastest ? (int)(gid % n) : (int)(gid / C);
  const int c = i_fastest ? (int)(gid / n) : (int)(gid % C);
  const double* x = s + (long long)i * si + (long long)c * sc;
  double sum = 0.0;  // t.sum / n, Trace.scala:69-71
  for (int t = 0; t < I; t++) sum += x[(long long)t * st];
  const double m = sum / (double)I;
  double ss = 0.0;  // t.map(a => pow(a - m, 2)).sum / (n - 1), Trace.scala:79-86
  for (int t = 0; t < I; t++) {
    const double d = x[(long long)t * st] - m;
    ss += d * d;
  }
  const size_t o = (size_t)i * C + c;
  mean[o] = m;
  var[o] = ss / (double)(I - 1);
  for (int lag = 1; lag <= max_lag; lag++) {  // Trace.variogram, Trace.scala:111-119
    double v = 0.0;
    for (int t = lag; t < I; t++) {
      const double d = x[(long long)t * st] - x[(long long)(t - lag) * st];
      v += d * d;
    }
    vario[(size_t)(lag - 1) * n * C + o] = v / (double)(I - lag);
  }
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
