// Probe: register layout of tcgen05.ld.16x256b.x8 (16 lanes x 64 columns per warp) and whether the
// lane base may be +16 inside a warp's 32-lane sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tsp_gnn_b200/csrc/tc_ptx.cuh"
using namespace tspgnn;

__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(256, 1) probe(int* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) ptx::tmem_alloc(&slot, 64);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const int q4 = warp & 3;
  const uint32_t t = slot + ((uint32_t)(q4 * 32) << 16);
  if (warp < 4) {     // fill: TMEM lane L (= q4*32 + lane), column c holds L*100 + c
    for (int c0 = 0; c0 < 64; c0 += 16) {
      float v[16];
      for (int j = 0; j < 16; ++j) v[j] = __int_as_float((q4 * 32 + lane) * 100 + c0 + j);
      ptx::tmem_st16(t + c0, v);
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  // warps 0-3 read lanes [32*q4, +16), warps 4-7 read lanes [32*q4+16, +16)
  const uint32_t base = slot + ((uint32_t)(q4 * 32 + (warp >> 2) * 16) << 16);
  uint32_t r[32];
  ld_16x256b_x8(base, r);
  for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 32 + i] = (int)r[i];
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(slot, 64);
}

int main() {
  int* d; cudaMalloc(&d, 256 * 32 * 4);
  probe<<<1, 256>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static int h[256 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  // check the hypothesis: thread t (t0 = t%4, t1 = t/4), register i = 4*k + 2*hi + j  ->
  // lane = base + t1 + 8*hi, column = 8*k + 2*t0 + j
  int bad = 0;
  for (int w = 0; w < 8; ++w)
    for (int t = 0; t < 32; ++t)
      for (int i = 0; i < 32; ++i) {
        const int k = i / 4, hi = (i / 2) % 2, j = i % 2, t0 = t % 4, t1 = t / 4;
        const int lane = (w & 3) * 32 + (w >> 2) * 16 + t1 + 8 * hi, col = 8 * k + 2 * t0 + j;
        if (h[(w * 32 + t) * 32 + i] != lane * 100 + col) ++bad;
      }
  printf("hypothesis mismatches: %d of %d\n", bad, 8 * 32 * 32);
  for (int w : {0, 5})
    for (int t : {0, 1, 5}) {
      printf("warp %d thread %2d:", w, t);
      for (int i = 0; i < 12; ++i) printf(" %5d", h[(w * 32 + t) * 32 + i]);
      printf("\n");
    }
  return 0;
}
