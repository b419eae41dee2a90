// Microbenchmark: cycles per warp-instruction per SMSP for the instruction kinds the epilogues use.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tsp_gnn_b200/csrc/tc_ptx.cuh"
using namespace tspgnn;

__constant__ float c_tab[4096];

template <int KIND>
__global__ void __launch_bounds__(512, 1) tput_kernel(int iters, long long* out, float* sink, int idx0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 x[8];
  for (int k = 0; k < 8; ++k) x[k] = make_float2(0.001f * (threadIdx.x + k), 1.0f + 0.002f * k);
  uint32_t u[8];
  for (int k = 0; k < 8; ++k) u[k] = threadIdx.x * 7 + k;
  float2 y[8], z[8];
  for (int k = 0; k < 8; ++k) { y[k] = make_float2(1.0f + 0.001f * k, 0.5f); z[k] = make_float2(0.9f, 1.0f - 0.001f * k); }
  __shared__ float4 sm[512];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    int idx = (idx0 + i * 16) & 1023;     // warp-uniform run-time index
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (KIND == 0) {          // indexed constant load (float2) feeding an FFMA2
        const float2 g = reinterpret_cast<const float2*>(c_tab + idx)[k];
        x[k] = __ffma2_rn(x[k], g, g);
      } else if (KIND == 1) {   // FFMA2 only
        x[k] = __ffma2_rn(x[k], x[(k + 1) & 7], x[(k + 2) & 7]);
      } else if (KIND == 2) {   // cvt.rn.bf16x2.f32
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[k]) : "f"(x[k].y), "f"(x[k].x));
        x[k].x = __uint_as_float(u[k]);
      } else if (KIND == 3) {   // FMNMX
        x[k].x = fmaxf(x[k].x, x[(k + 1) & 7].y);
        x[k].y = fminf(x[k].y, x[(k + 3) & 7].x);
      } else if (KIND == 4) {   // MUFU.RCP
        x[k].x = ptx::rcp_approx(x[k].x);
      } else if (KIND == 5) {   // STS.128 + fence.proxy.async (latency of the fence)
        sm[threadIdx.x] = make_float4(x[k].x, x[k].y, 0.f, 1.f);
        ptx::fence_proxy_async_smem();
      } else if (KIND == 6) {   // LDS.128 broadcast (all lanes same address)
        const float4 g = sm[(idx + k) & 511];
        x[k] = __ffma2_rn(x[k], make_float2(g.x, g.y), make_float2(g.z, g.w));
      } else if (KIND == 8) {   // MUFU.EX2 and two independent FFMA2 per slot (do the pipes overlap?)
        x[k].x = ptx::ex2_approx(x[k].x);
        y[k] = __ffma2_rn(y[k], y[(k + 1) & 7], y[(k + 2) & 7]);
        z[k] = __ffma2_rn(z[k], z[(k + 1) & 7], z[(k + 2) & 7]);
      } else if (KIND == 9) {   // MUFU.EX2 + 4 scalar FMNMX (alu pipe)
        x[k].x = ptx::ex2_approx(x[k].x);
        y[k].x = fmaxf(y[k].x, y[(k + 1) & 7].y);
        y[k].y = fminf(y[k].y, y[(k + 3) & 7].x);
        z[k].x = fmaxf(z[k].x, z[(k + 1) & 7].y);
        z[k].y = fminf(z[k].y, z[(k + 3) & 7].x);
      } else if (KIND == 10) {  // FFMA2 + 2 FMNMX (fma pipe vs alu pipe)
        y[k] = __ffma2_rn(y[k], y[(k + 1) & 7], y[(k + 2) & 7]);
        z[k].x = fmaxf(z[k].x, z[(k + 1) & 7].y);
        z[k].y = fminf(z[k].y, z[(k + 3) & 7].x);
      } else if (KIND == 11) {  // two scalar FFMA, register operands
        x[k].x = fmaf(x[k].x, x[(k + 1) & 7].x, x[(k + 2) & 7].x);
        x[k].y = fmaf(x[k].y, x[(k + 1) & 7].y, x[(k + 2) & 7].y);
      } else if (KIND == 12) {  // FMUL2 / FADD2 (two register pairs)
        x[k] = __fmul2_rn(x[k], y[k]);
        z[k] = __fadd2_rn(z[k], y[(k + 1) & 7]);
      } else if (KIND == 7) {   // scalar FFMA with constant operand c[][] immediate address
        x[k].x = fmaf(x[k].x, c_tab[k], c_tab[k + 8]);
        x[k].y = fmaf(x[k].y, c_tab[k + 16], c_tab[k + 24]);
      }
    }
  }
  const long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  float s = 0;
  for (int k = 0; k < 8; ++k) s += x[k].x + x[k].y + u[k] + y[k].x + y[k].y + z[k].x + z[k].y;
  if (s == 123.456f) sink[0] = s + sm[0].x;
}

template <int KIND>
void run(const char* name, long long* d, float* sink, double per_iter) {
  long long h[16];
  for (int nthreads : {128, 256, 512}) {
    const int iters = 1000;
    tput_kernel<KIND><<<1, nthreads>>>(iters, d, sink, 4);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const int wps = nthreads / 128;   // warps per SMSP
    printf("%-28s warps/SMSP=%d  cycles per warp-instr per SMSP = %6.2f   (%s)\n", name, wps,
           (double)h[0] / (iters * per_iter) / wps, cudaGetErrorString(e));
  }
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 16 * 8); cudaMalloc(&sink, 4);
  float tab[4096]; for (int i = 0; i < 4096; ++i) tab[i] = 1.0f + 1e-4f * i;
  cudaMemcpyToSymbol(c_tab, tab, sizeof(tab));
  run<0>("LDC.64 indexed + FFMA2", d, sink, 8);
  run<1>("FFMA2", d, sink, 8);
  run<2>("cvt.rn.bf16x2 (F2FP)", d, sink, 8);
  run<3>("FMNMX x2", d, sink, 16);
  run<4>("MUFU.RCP", d, sink, 8);
  run<5>("STS.128 + fence.proxy.async", d, sink, 8);
  run<6>("LDS.128 bcast + FFMA2", d, sink, 8);
  run<7>("FFMA x2 const-operand", d, sink, 16);
  run<11>("FFMA x2 scalar regs", d, sink, 16);
  run<12>("FMUL2 + FADD2", d, sink, 16);
  run<8>("slot: MUFU + 2 FFMA2", d, sink, 8);
  run<9>("slot: MUFU + 4 FMNMX", d, sink, 8);
  run<10>("slot: FFMA2 + 2 FMNMX", d, sink, 8);
  return 0;
}
