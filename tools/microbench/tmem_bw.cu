// Microbenchmark: TMEM read bandwidth seen by tcgen05.ld (32x32b.x32) with 4 or 8 warps per SM,
// and MUFU throughput.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tsp_gnn_b200/csrc/tc_ptx.cuh"
using namespace tspgnn;

__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, int ncols_per_ld, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) ptx::tmem_alloc(&slot, 512);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t t = slot + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 256;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (ncols_per_ld == 64) {
      float v[64];
      ptx::tmem_ld64(t + (i & 3) * 64, v);
#pragma unroll
      for (int j = 0; j < 64; j += 16) acc += v[j];
    } else {
      float v[16];
      ptx::tmem_ld16(t + (i & 15) * 16, v);
      acc += v[0] + v[8];
    }
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(slot, 512);
}

__global__ void __launch_bounds__(512, 1) mufu_kernel(int iters, long long* out, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float x[8];
  for (int k = 0; k < 8; ++k) x[k] = 0.001f * (threadIdx.x + k);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = ptx::ex2_approx(x[k]);
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  float s = 0; for (int k = 0; k < 8; ++k) s += x[k];
  if (s == 123.456f) sink[0] = s;
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 148 * 16 * 8); cudaMalloc(&sink, 4);
  long long h[16];
  for (int nthreads : {128, 256, 512}) {
    for (int cols : {64, 16}) {
      const int iters = 2000;
      tmem_read_kernel<<<1, nthreads>>>(iters, cols, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double bytes = (double)iters * cols * 4 * 32;   // per warp
      printf("tmem ld x%d  warps=%2d  cycles/ld(warp0)=%.1f  B/cycle/warp=%.1f  B/cycle/SM=%.1f  (%s)\n", cols, nthreads / 32,
             (double)h[0] / iters, bytes / h[0], bytes * (nthreads / 32) / h[0], cudaGetErrorString(e));
    }
  }
  for (int nthreads : {128, 256, 512}) {
    const int iters = 2000;
    mufu_kernel<<<1, nthreads>>>(iters, d, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mufu ex2 warps=%2d cycles per warp-instr = %.2f  (per SMSP: %.2f)\n", nthreads / 32, (double)h[0] / (iters * 8.0),
           (double)h[0] / (iters * 8.0) / ((nthreads / 32 + 3) / 4));
  }
  return 0;
}
