// Microbenchmark: how fast can one SM push 256-byte row reductions into an L2-resident table?
//   RED4 : red.global.add.v4.f32, a half-warp covers one row (the scatter of K2 as first built)
//   BULK : cp.reduce.async.bulk.global.shared::cta.add.f32, one instruction per row issued by one lane
//          (the TMA engine performs the reduction; the LSU only sees one instruction per 32 rows)
// Every CTA (one per SM) reduces rows of its shared-memory tile into pseudo-random rows of a 5120 x 64
// fp32 table (the xV buffer of the north-star batch), like 148 edge CTAs do at once.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_red bulk_red.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tsp_gnn_b200/csrc/tc_ptx.cuh"
using namespace tspgnn;

__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// mode 0: RED.v4 (nwarps warps, 2 rows per instruction); mode 1: bulk reduce (nwarps warps, 32 rows per instruction)
__global__ void __launch_bounds__(512, 1) red_kernel(int mode, int nwarps, int iters, int row_stride, float* table, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* tile = reinterpret_cast<float*>(smem);
  for (int i = tid; i < 128 * row_stride / 4; i += blockDim.x) tile[i] = 1.0f;
  __shared__ int dst_tab[780];
  for (int eidx = tid; eidx < 780; eidx += blockDim.x) {
    int si = 0, rem = eidx;
    while (rem >= 39 - si) { rem -= 39 - si; ++si; }
    dst_tab[eidx] = si + 1 + rem;
  }
  ptx::fence_proxy_async_smem();
  __syncthreads();
  const uint32_t tile_s = ptx::smem_u32(tile);
  uint32_t h = blockIdx.x * 7919u + warp * 104729u + 17u;
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      h = h * 1664525u + 1013904223u;
      const uint32_t base = (h >> 10) % 5000u;          // 32 consecutive-ish target rows like dst of a run
      if (mode == 0) {
        const int hw = lane >> 4, c16 = lane & 15;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int rr = 2 * i + hw;
          const float4 m = ptx::lds128f(tile_s + ((warp & 3) * 32 + rr) * row_stride + c16 * 16);
          ptx::red_add_v4(table + static_cast<int64_t>(base + rr) * 64 + 4 * c16, m);
        }
      } else if (mode == 4) {
        // real pattern: the 128 edge rows of tile (it % 6) of a 40-vertex complete graph owned by this CTA;
        // dst of row r of the instance's sorted edge list, every warp its own 32 rows
        const int hw = lane >> 4, c16 = lane & 15;
        const uint32_t vbase = (blockIdx.x * 40u) % 5080u;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int rr = 2 * i + hw;
          int eidx = (it % 6) * 128 + (warp & 3) * 32 + rr;     // edge index inside the instance (780 edges)
          if (eidx >= 780) eidx -= 780;
          const int di = dst_tab[eidx];
          const float4 m = ptx::lds128f(tile_s + ((warp & 3) * 32 + rr) * row_stride + c16 * 16);
          ptx::red_add_v4(table + static_cast<int64_t>(vbase + di) * 64 + 4 * c16, m);
        }
      } else if (mode == 5 || mode == 6) {
        // mode 5: half-warps diverged, each issues its own 16-lane RED (one row per instruction)
        // mode 6: 8 active lanes (half a row per instruction)
        const int hw = lane >> 4, c16 = lane & 15;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int rr = 2 * i + hw;
          const float4 m = ptx::lds128f(tile_s + ((warp & 3) * 32 + rr) * row_stride + c16 * 16);
          if (mode == 5) {
            if (hw == 0) ptx::red_add_v4(table + static_cast<int64_t>(base + rr) * 64 + 4 * c16, m);
            __syncwarp();
            if (hw == 1) ptx::red_add_v4(table + static_cast<int64_t>(base + rr) * 64 + 4 * c16, m);
            __syncwarp();
          } else {
            if ((lane & 8) == 0) ptx::red_add_v4(table + static_cast<int64_t>(base + rr) * 64 + 4 * c16, m);
          }
        }
      } else if (mode == 2) {
        // thread = row: every lane reduces the 16 chunks of its own row (32 distinct rows per instruction)
        float* rowp = table + static_cast<int64_t>(base + lane) * 64;
#pragma unroll
        for (int q = 0; q < 16; ++q) ptx::red_add_v4(rowp + 4 * q, make_float4(1.f, 2.f, 3.f, 4.f));
      } else if (mode == 3) {
        // thread = row, but lane pairs cover a 32-byte sector: lane l handles chunk 2q + (l & 1) of row (l >> 1) (+16)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          float* rowp = table + static_cast<int64_t>(base + (lane >> 1) + 16 * hrow) * 64 + 4 * (lane & 1);
#pragma unroll
          for (int q = 0; q < 8; ++q) ptx::red_add_v4(rowp + 8 * q, make_float4(1.f, 2.f, 3.f, 4.f));
        }
      } else {
        bulk_reduce_add_f32(table + static_cast<int64_t>(base + lane) * 64, tile_s + ((warp & 3) * 32 + lane) * row_stride, 256);
        bulk_commit();
        if ((it & 7) == 7) bulk_wait_read0();
      }
    }
    if (mode == 1) bulk_wait0();
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
}

int main() {
  float* table;
  long long* d_out;
  cudaMalloc(&table, 5120 * 64 * 4);
  cudaMemset(table, 0, 5120 * 64 * 4);
  cudaMalloc(&d_out, 148 * 16 * 8);
  const int smem = 128 * 272;
  cudaFuncSetAttribute(red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  static long long h[148 * 16];
  for (int grid : {1, 148}) {
    for (int mode : {0, 5, 6}) {
      for (int nwarps : {1, 4, 12}) {
        for (int stride : {256, 272}) {
          if (stride == 256) continue;
          const int iters = 400;
          red_kernel<<<grid, 512, smem>>>(mode, nwarps, iters, stride, table, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          cudaMemcpy(h, d_out, sizeof(long long) * grid * 16, cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int b = 0; b < grid; ++b)
            for (int w = 0; w < nwarps; ++w) mx = h[b * 16 + w] > mx ? h[b * 16 + w] : mx;
          const double rows = (double)iters * 32 * nwarps;
          printf("grid=%3d %-5s warps=%2d stride=%d : %8.2f cycles per 256-B row per SM  (%6.1f B/clk/SM)  %s\n", grid,
                 mode == 0 ? "RED4" : (mode == 2 ? "ROW" : (mode == 3 ? "PAIR" : (mode == 4 ? "REAL" : (mode == 5 ? "HALF" : (mode == 6 ? "QUART" : "BULK"))))), nwarps, stride, mx / rows, rows * 256 / mx, cudaGetErrorString(e));
        }
      }
    }
  }
  // correctness of the bulk reduction: table must hold integer counts, total = rows reduced
  return 0;
}
