// Probe of the CTA-pair protocol the fused timestep kernel relies on (cta_group::2):
//   * both CTAs of a 2-CTA cluster write their own A rows (128 x 64 bf16) and their half of the B
//     image (N/2 output features x 64 k) with ordinary shared-memory stores,
//   * the odd CTA tells the even one "my operands are visible" with a remote mbarrier arrive,
//   * the even CTA issues tcgen05.mma.cta_group::2 (M = 256) and commits to the barrier of BOTH CTAs,
//   * each CTA reads its 128 rows x N columns back from its own TMEM.
// Checks D = A . B^T against the host for N = 256 (LSTM shape) and N = 64 (MLP layer shape).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cta2_mma cta2_mma.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../tsp_gnn_b200/csrc/tc_ptx.cuh"
using namespace tspgnn;

static uint16_t bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// A: [256 rows][64 k] bf16 row-major in global; B: [N][64 k] bf16 row-major (output feature n, k)
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
pair_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, float* __restrict__ Dout, int reps,
            long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* a_sm = smem;                       // 128 x 64 bf16, chunk-major: (k/8)*2048 + r*16 + (k%8)*2
  uint8_t* b_sm = smem + 16384;               // (N/2) x 64 bf16: (k/8)*(N/2*16) + n*16 + (k%8)*2
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + (N / 2) * 128);
  uint64_t* peer_full = bars;                 // leader: arrived by the odd CTA
  uint64_t* acc_full = bars + 1;              // both: committed by the MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  if (tid == 0) {
    ptx::mbar_init(peer_full, 1);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) ptx::tmem_alloc_pair(tmem_slot, 256);
  // operands: this CTA's rows of A, this CTA's half of B
  for (int i = tid; i < 128 * 64; i += blockDim.x) {
    const int r = i >> 6, k = i & 63;
    *reinterpret_cast<uint16_t*>(a_sm + (k >> 3) * 2048 + r * 16 + (k & 7) * 2) = A[(rank * 128 + r) * 64 + k];
  }
  for (int i = tid; i < (N / 2) * 64; i += blockDim.x) {
    const int n = i >> 6, k = i & 63;
    *reinterpret_cast<uint16_t*>(b_sm + (k >> 3) * ((N / 2) * 16) + n * 16 + (k & 7) * 2) =
        B[(rank * (N / 2) + n) * 64 + k];
  }
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  ptx::cluster_sync();      // barriers of both CTAs initialised, both TMEM allocations done
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(256, N);
  long long t0 = 0, t1 = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (warp == 4) {
      if (rank == 1) {
        // relay: my operands are written (the __syncthreads / previous acc_full ordered them)
        if (lane == 0) ptx::mbar_arrive_remote(ptx::mapa_u32(ptx::smem_u32(peer_full), 0));
        __syncwarp();
      } else {
        ptx::mbar_wait_cluster(peer_full, rep & 1);
        ptx::tcgen05_fence_after();
        if (rep == 1) t0 = clock64();
        if (ptx::elect_one()) {
          const uint64_t adesc = ptx::umma_desc_k_nosw(ptx::smem_u32(a_sm), 2048, 128);
          const uint64_t bdesc = ptx::umma_desc_k_nosw(ptx::smem_u32(b_sm), (N / 2) * 16, 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss_pair(tmem, adesc + ((k * 4096) >> 4), bdesc + ((k * 2 * (N / 2) * 16) >> 4), IDESC,
                                   k ? 1u : 0u);
          ptx::umma_commit_pair(acc_full);
        }
        __syncwarp();
      }
    }
    // everybody (both CTAs) waits for the accumulator
    ptx::mbar_wait(acc_full, rep & 1);
    ptx::tcgen05_fence_after();
    if (warp == 4 && rank == 0 && rep == reps - 1) t1 = clock64();
    if (warp < 4 && rep == reps - 1) {
      const int r = warp * 32 + lane;
      const uint32_t t_acc = tmem + (static_cast<uint32_t>(warp * 32) << 16);
      for (int c0 = 0; c0 < N; c0 += 64) {
        float v[64];
        ptx::tmem_ld64(t_acc + c0, v);
        for (int j = 0; j < 64; ++j) Dout[(size_t)(rank * 128 + r) * N + c0 + j] = v[j];
      }
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
  }
  if (warp == 4 && rank == 0 && lane == 0 && cycles) cycles[0] = t1 - t0;
  ptx::tcgen05_fence_before();
  ptx::cluster_sync();      // the peer's shared memory / TMEM stay alive until every MMA has completed
  if (warp == 4) ptx::tmem_dealloc_pair(tmem, 256);
}

template <int N>
static int run(const char* name) {
  std::vector<uint16_t> A(256 * 64), B(N * 64);
  srand(7 + N);
  for (auto& x : A) x = bf16_rn((rand() % 2001 - 1000) / 1000.0f);
  for (auto& x : B) x = bf16_rn((rand() % 2001 - 1000) / 4000.0f);
  uint16_t *dA, *dB;
  float* dD;
  long long* dC;
  cudaMalloc(&dA, A.size() * 2);
  cudaMalloc(&dB, B.size() * 2);
  cudaMalloc(&dD, 256 * N * 4);
  cudaMalloc(&dC, 8);
  cudaMemset(dD, 0xff, 256 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 16384 + (N / 2) * 128 + 64 + 128;
  cudaFuncSetAttribute(pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 9;
  pair_kernel<N><<<2, 192, smem>>>(dA, dB, dD, reps, dC);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s: launch status: %s\n", name, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> D(256 * N);
  long long cyc = 0;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)bf16_f(A[m * 64 + k]) * bf16_f(B[n * 64 + k]);
      const double err = fabs(s - D[m * N + n]);
      if (!(err < 1e-3)) ++bad;
      if (err > maxerr || err != err) maxerr = err;
    }
  printf("%s: M=256 N=%d K=64 cta_group::2  max|err| = %.3e  mismatches = %d of %d  (%lld cycles for %d rounds)\n", name, N,
         maxerr, bad, 256 * N, cyc, reps - 2);
  if (bad) {
    for (int m : {0, 1, 127, 128, 255})
      printf("  row %3d: got %9.5f %9.5f ... %9.5f\n", m, D[m * N], D[m * N + 1], D[m * N + N - 1]);
  }
  return bad != 0;
}

int main() {
  int rc = run<256>("lstm-shape");
  rc |= run<64>("mlp-shape");
  printf(rc ? "CTA2 PROBE FAILED\n" : "CTA2 PROBE OK\n");
  return rc;
}
