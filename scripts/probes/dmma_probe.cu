// scripts/probes/dmma_probe.cu -- measurement probe, NOT part of the product.
//
// Question it answers on a B200 before the chain-batched contraction is built: (1) is the operand mapping of
// mma.sync.aligned.m8n8k4.row.col.f64 what DESIGN.md assumes, on the data layout the sampler already uses (tile-major
// observation columns [tile][column][32 rows]; parameters [d][chains], chain fastest); (2) what rate does the fp64
// tensor-core path sustain for Z[rows x chains] = X[rows x d] * B[d x chains] at cfg 3's shape (100 000 rows, d = 50 padded
// to 52, 2048 chains) against the plain-DFMA form of the same product, both reading operands through L1/L2.
//
//   fragments (PTX ISA, m8n8k4 .f64):  A 8x4 row-major : lane -> A[lane / 4][lane % 4]
//                                      B 4x8 col-major : lane -> B[lane % 4][lane / 4]
//                                      C/D 8x8         : lane -> C[lane / 4][2 * (lane % 4) + {0, 1}]
//
// Build (cross-compiles without a GPU): nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
//        scripts/probes/dmma_probe.cu -o build/libdmma_probe.so        (scripts/probe_dmma.py does this and runs it)
#include <cuda_runtime.h>

#include <cstdio>

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// X: tile-major [n_tiles][dpad][32]; B: [dpad][chains]; Z: [n_tiles*32][chains] (row-major).  One warp: 32 rows x 8 chains.
extern "C" __global__ void k_dmma(const double* __restrict__ X, const double* __restrict__ B, double* __restrict__ Z, int n_tiles, int dpad,
                                  int chains) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x, cbase = (blockIdx.y * (blockDim.x >> 5) + warp) * 8;
  if (tile >= n_tiles || cbase >= chains) return;
  const double* xt = X + (size_t)tile * dpad * 32;
  double acc[4][2] = {};
  for (int kb = 0; kb < dpad; kb += 4) {
    const double b = __ldg(B + (size_t)(kb + (lane & 3)) * chains + cbase + (lane >> 2));
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {  // 4 row blocks of 8 rows share the B fragment
      const double a = __ldg(xt + (size_t)(kb + (lane & 3)) * 32 + mt * 8 + (lane >> 2));
      dmma_m8n8k4(acc[mt][0], acc[mt][1], a, b);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    double* z = Z + (size_t)(tile * 32 + mt * 8 + (lane >> 2)) * chains + cbase + 2 * (lane & 3);
    z[0] = acc[mt][0];
    z[1] = acc[mt][1];
  }
}

// the same product the rows-across-lanes way: lane = row, one chain per warp, sequential DFMA over the d columns
extern "C" __global__ void k_dfma(const double* __restrict__ X, const double* __restrict__ B, double* __restrict__ Z, int n_tiles, int dpad,
                                  int chains) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x, c = blockIdx.y * (blockDim.x >> 5) + warp;
  if (tile >= n_tiles || c >= chains) return;
  const double* xt = X + (size_t)tile * dpad * 32;
  double z = 0.0;
  for (int j = 0; j < dpad; j++) z = fma(__ldg(B + (size_t)j * chains + c), __ldg(xt + (size_t)j * 32 + lane), z);
  Z[(size_t)(tile * 32 + lane) * chains + c] = z;
}

// which: 0 = DMMA, 1 = DFMA.  Returns the mean kernel time in ms over `iters` launches (after one warm-up), < 0 on error.
extern "C" float dmma_probe_run(int which, const double* dX, const double* dB, double* dZ, int n_tiles, int dpad, int chains, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int wpb = 4;
  dim3 block(32 * wpb), grid(n_tiles, which == 0 ? (chains / 8 + wpb - 1) / wpb : (chains + wpb - 1) / wpb);
  for (int k = -1; k < iters; k++) {
    if (k == 0) cudaEventRecord(e0);
    if (which == 0)
      k_dmma<<<grid, block>>>(dX, dB, dZ, n_tiles, dpad, chains);
    else
      k_dfma<<<grid, block>>>(dX, dB, dZ, n_tiles, dpad, chains);
  }
  cudaEventRecord(e1);
  if (cudaEventSynchronize(e1) != cudaSuccess) return -1.0f;
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return cudaGetLastError() == cudaSuccess ? ms / iters : -1.0f;
}
