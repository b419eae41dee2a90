// Microbenchmark: who shares the SM's L1 / shared-memory data pipe?
// One CTA on one SM; every warp gets a role and runs it for a fixed window of cycles, then reports how
// many operations it completed.  Roles:
//   M256 / M64 : one warp issues tcgen05.mma (SS, M128 x N256 or N64, K16, bf16) in batches of 8 with
//                two commit barriers (the tensor pipe never drains)
//   LDSD       : LDS.128, every lane its own 16 bytes (4 wavefronts of 128 B per instruction)
//   LDSB       : LDS.128, all lanes the same address (the LayerNorm-parameter loads of the epilogue)
//   STSD       : STS.128 distinct addresses
//   TLD        : tcgen05.ld 32x32b.x32 x2 (64 columns = 8 KB per warp instruction pair)
//   TST        : tcgen05.st 32x32b.x16
//   LDG        : LDG.256 from a 1.3 MB L2-resident table, 256-B rows (the gather of K1)
// Output: operations and bytes per cycle of every role in every scenario.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o datapipe datapipe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <cuda_runtime.h>
#include "../../tsp_gnn_b200/csrc/tc_ptx.cuh"
using namespace tspgnn;

enum Role : int { IDLE = 0, M256, M64, LDSD, LDSB, STSD, TLD, TST, LDG, FMA, M64SW, M64TS, M256TS, LDCU, LDSB64, M128, M64MN, M256MN, NROLE };
__constant__ float2 c_tab[1024];
struct Cfg { int role[16]; long long window; };

__global__ void __launch_bounds__(512, 1) pipe_kernel(Cfg cfg, const float* __restrict__ table, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* a_sm = smem;               // 128 x 64 bf16 = 16 KB
  uint8_t* b_sm = smem + 16384;       // 256 x 64 bf16 = 32 KB
  uint8_t* scratch = smem + 49152;    // 64 KB for the LDS / STS roles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152 + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (49152 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int role = cfg.role[warp];
  const long long t0 = clock64();
  long long ops = 0;
  float acc = 0.f;
  if (role == M256 || role == M64) {
    const int N = (role == M256) ? 256 : 64;
    const uint32_t idesc = (role == M256) ? ptx::umma_idesc_bf16(128, 256) : ptx::umma_idesc_bf16(128, 64);
    const uint64_t adesc = ptx::umma_desc_k_nosw(ptx::smem_u32(a_sm), 2048, 128);
    const uint64_t bdesc = ptx::umma_desc_k_nosw(ptx::smem_u32(b_sm), N * 16, 128);
    int batch = 0;
    while (clock64() - t0 < cfg.window) {
      if (batch >= 2) ptx::mbar_wait(&bars[batch & 1], ((batch >> 1) - 1) & 1);
      if (ptx::elect_one()) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss(tmem, adesc + ((k * 4096) >> 4), bdesc + ((k * 2 * N * 16) >> 4), idesc, 1u);
        ptx::umma_commit(&bars[batch & 1]);
      }
      __syncwarp();
      ++batch;
      ops += 8;
    }
    // drain
    if (batch >= 1) ptx::mbar_wait(&bars[(batch - 1) & 1], ((batch - 1) >> 1) & 1);
    if (batch >= 2) ptx::mbar_wait(&bars[(batch - 2) & 1], ((batch - 2) >> 1) & 1);
  } else if (role == LDSD || role == LDSB) {
    const uint32_t base = ptx::smem_u32(scratch) + (role == LDSD ? ((warp & 7) * 4096 + lane * 16) : ((warp & 7) * 4096));
    uint32_t off = 0;
    while (clock64() - t0 < cfg.window) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 v = ptx::lds128f(base + ((off + u * 512) & 3583));
        acc += v.x;
      }
      off += 16;
      ops += 8;
    }
  } else if (role == STSD) {
    const uint32_t base = ptx::smem_u32(scratch) + (warp & 7) * 4096 + lane * 16;
    uint32_t off = 0;
    while (clock64() - t0 < cfg.window) {
#pragma unroll
      for (int u = 0; u < 8; ++u) ptx::sts128(base + ((off + u * 512) & 3583), make_uint4(off, u, lane, warp));
      off += 16;
      ops += 8;
    }
  } else if (role == TLD) {
    const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    int i = 0;
    while (clock64() - t0 < cfg.window) {
      float v[64];
      ptx::tmem_ld64(t + (i & 3) * 64, v);
#pragma unroll
      for (int j = 0; j < 64; j += 16) acc += v[j];
      ++i;
      ops += 1;
    }
  } else if (role == TST) {
    const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    int i = 0;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = lane + j;
    while (clock64() - t0 < cfg.window) {
      ptx::tmem_st16(t + (i & 15) * 16, v);
      ++i;
      ops += 1;
    }
  } else if (role == LDG) {
    // 5120 rows of 256 B; a warp instruction covers 8 rows x 4 chunks of 32 B like the K1 gather
    uint32_t h = warp * 977 + 13;
    while (clock64() - t0 < cfg.window) {
      float u[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h = h * 1664525u + 1013904223u;
        const uint32_t row = ((h >> 8) + (lane & 7) * 131) % 5120u;
        ptx::ldg256(table + row * 64 + (lane >> 3) * 8 + (q & 1) * 32, u[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) acc += u[q][0] + u[q][7];
      ops += 4;
    }

  } else if (role == M64SW || role == M64TS || role == M256TS || role == M128 || role == M64MN || role == M256MN) {
    const int N = (role == M256TS || role == M256MN) ? 256 : (role == M128 ? 128 : 64);
    const uint32_t idesc = (N == 256) ? ptx::umma_idesc_bf16(128, 256) : (N == 128 ? ptx::umma_idesc_bf16(128, 128) : ptx::umma_idesc_bf16(128, 64));
    uint64_t adesc, bdesc;
    uint32_t kstepA, kstepB;
    if (role == M64SW) {
      // SWIZZLE_128B K-major: rows of 128 B (64 bf16), 8-row groups 1024 B apart
      const uint32_t a_al = (ptx::smem_u32(a_sm) + 1023) & ~1023u, b_al = (ptx::smem_u32(b_sm) + 1023) & ~1023u;
      adesc = static_cast<uint64_t>((a_al & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
      bdesc = static_cast<uint64_t>((b_al & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
      kstepA = 32 >> 4;
      kstepB = 32 >> 4;
    } else {
      adesc = ptx::umma_desc_k_nosw(ptx::smem_u32(a_sm), 2048, 128);
      bdesc = ptx::umma_desc_k_nosw(ptx::smem_u32(b_sm), N * 16, 128);
      kstepA = 4096 >> 4;
      kstepB = (2 * N * 16) >> 4;
    }
    uint32_t idesc2 = idesc;
    if (role == M64MN || role == M256MN) {      // both operands MN-major: the chunk-major image read as [col][row]
      adesc = ptx::umma_desc_k_nosw(ptx::smem_u32(a_sm), 128, 2048);
      bdesc = ptx::umma_desc_k_nosw(ptx::smem_u32(b_sm), 128, 2048);
      kstepA = kstepB = 256 >> 4;
      idesc2 |= (1u << 15) | (1u << 16);
    }
    const bool ts = (role == M64TS || role == M256TS);
    const uint32_t a_tmem = tmem + 448;     // A operand in TMEM: 4 k-steps x 8 columns
    int batch = 0;
    while (clock64() - t0 < cfg.window) {
      if (batch >= 2) ptx::mbar_wait(&bars[batch & 1], ((batch >> 1) - 1) & 1);
      if (ptx::elect_one()) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (ts) {
              asm volatile(
                  "{\n\t.reg .pred p;\n\t"
                  "setp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
                  "r"(a_tmem + k * 8), "l"(bdesc + k * kstepB), "r"(idesc), "r"(1u)
                  : "memory");
            } else {
              ptx::umma_bf16_ss(tmem, adesc + k * kstepA, bdesc + k * kstepB, idesc2, 1u);
            }
          }
        ptx::umma_commit(&bars[batch & 1]);
      }
      __syncwarp();
      ++batch;
      ops += 8;
    }
    if (batch >= 1) ptx::mbar_wait(&bars[(batch - 1) & 1], ((batch - 1) >> 1) & 1);
    if (batch >= 2) ptx::mbar_wait(&bars[(batch - 2) & 1], ((batch - 2) >> 1) & 1);
  } else if (role == LDCU) {
    int i = 0;
    while (clock64() - t0 < cfg.window) {
      const int base = (i & 63) * 8;      // warp-uniform run-time index
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float2 v = c_tab[base + u + (i >> 20)];
        acc += v.x * v.y;
      }
      ++i;
      ops += 8;
    }
  } else if (role == LDSB64) {
    const uint32_t base = ptx::smem_u32(scratch) + (warp & 7) * 4096;
    uint32_t off = 0;
    while (clock64() - t0 < cfg.window) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(base + ((off + u * 512) & 3583)));
        acc += v.x;
      }
      off += 16;
      ops += 8;
    }
  } else if (role == FMA) {
    float2 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = make_float2(0.001f * lane + k, 0.5f);
    while (clock64() - t0 < cfg.window) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = __ffma2_rn(x[k], make_float2(0.999f, 1.001f), make_float2(0.25f, 0.125f));
      ops += 32;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += x[k].x + x[k].y;
  }
  const long long t1 = clock64();
  if (lane == 0) {
    out[warp * 2] = ops;
    out[warp * 2 + 1] = t1 - t0;
  }
  if (acc == 123.456f) sink[0] = acc;
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

static const char* rname[] = {"idle", "M256", "M64", "LDSD", "LDSB", "STSD", "TLD", "TST", "LDG", "FMA", "M64SW", "M64TS", "M256TS", "LDCU", "LDSB64", "M128", "M64MN", "M256MN"};
static double bytes_per_op(int r) {
  switch (r) {
    case M256: return 4096 + 8192;
    case M64: case M64SW: case M64MN: return 4096 + 2048;
    case M256MN: return 4096 + 8192;
    case M64TS: return 2048;
    case M256TS: return 8192;
    case M128: return 8192;
    case LDCU: case LDSB64: return 8;
    case LDSD: case STSD: return 512;
    case LDSB: return 16;
    case TLD: return 8192;
    case TST: return 2048;
    case LDG: return 1024;
    default: return 0;
  }
}

int main() {
  float* table;
  long long* d_out;
  float* sink;
  cudaMalloc(&table, 5120 * 64 * 4);
  cudaMemset(table, 0, 5120 * 64 * 4);
  cudaMalloc(&d_out, 16 * 2 * 8);
  cudaMalloc(&sink, 4);
  const int smem = 49152 + 65536 + 64 + 128;
  cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Sc { const char* name; std::vector<std::pair<int, int>> roles; };   // (role, number of warps); warp 0 is the first entry
  std::vector<Sc> scs = {
      {"mma N64 K-major alone", {{M64, 1}}},
      {"mma N64 MN-major alone", {{M64MN, 1}}},
      {"mma N256 MN-major alone", {{M256MN, 1}}},
      {"mma N256 MN-major + sts x8", {{M256MN, 1}, {IDLE, 3}, {STSD, 8}}},
  };
  for (auto& sc : scs) {
    Cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.window = 400000;
    int w = 0;
    for (auto& pr : sc.roles)
      for (int i = 0; i < pr.second && w < 16; ++i) cfg.role[w++] = pr.first;
    pipe_kernel<<<1, 512, smem>>>(cfg, table, d_out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("== %-46s (%s)\n", sc.name, cudaGetErrorString(e));
    double tot[NROLE] = {0};
    double cyc[NROLE] = {0};
    int cnt[NROLE] = {0};
    for (int i = 0; i < 16; ++i) {
      const int r = cfg.role[i];
      if (r == IDLE) continue;
      tot[r] += (double)h[2 * i];
      cyc[r] += (double)h[2 * i + 1];
      cnt[r]++;
    }
    double allbytes = 0;
    for (int r = 1; r < NROLE; ++r)
      if (cnt[r]) {
        const double c = cyc[r] / cnt[r];
        printf("     %-5s warps=%2d  ops/SM=%9.0f  cycles/op(SM)=%8.2f  B/cycle(SM)=%7.1f\n", rname[r], cnt[r], tot[r], c / tot[r],
               tot[r] * bytes_per_op(r) / c);
        allbytes += tot[r] * bytes_per_op(r) / c;
      }
    printf("     sum B/cycle = %.1f\n", allbytes);
  }
  return 0;
}
