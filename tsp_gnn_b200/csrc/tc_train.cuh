// Tensor-core (tcgen05) row GEMM of the reverse pass (model.py:157-167 through tf.gradients):
//   Y[rows, 64 NB] = epi( X[rows, 64 KB] . Wm (+ bias) ),   Wm = W or W^T of a parameter matrix
// for row-major fp32 matrices addressed as (pointer, leading dimension) per 64-column block -- the same
// contract as rowgemm_kernel (train_kernels.cuh), which stays the fp32 CUDA-core form of the `simt` mode.
// Used for the recompute of the LSTM pre-activations z = [x, h] . K (KB 2, NB 4), for [dx, dh] = dz . K^T
// (KB 4, NB 2) and for the 64-wide layers of the message / vote / initial-embedding MLPs, forward (recompute,
// bias + ReLU) and reverse (W^T, ReLU mask or accumulation).
//
// Arithmetic: the bf16x3 scheme of the forward kernels (operands split into bf16 hi + lo, three MMAs per
// product, fp32 accumulation in TMEM), so the reverse pass sees the same ~2^-16 relative operand error as
// the forward pass whose state it differentiates.
//
// Structure (one persistent CTA per SM, 512 threads, tiles of 128 rows):
//   warps 0-3, 4-7 : two epilogue warpgroups, one tile each.  The accumulator is read in the 16-lane
//                    "quad" TMEM shape, so the four threads of a quad hold 32 contiguous bytes of a row and
//                    the fp32 row-major output goes to global memory in full sectors without a shared-memory
//                    transposition (8 rows x 32 bytes per warp instruction).
//   warp  8        : tcgen05.mma issuer, owns the TMEM allocation (two accumulators of 64 NB columns)
//   warps 9-15     : producers: 128 rows x 64 columns of X per k-block, fp32 -> bf16 hi / lo planes in the
//                    un-swizzled K-major canonical layout (chunk-major, as the forward kernels' tile images),
//                    through a ring of 32 KB slots
// The B operand (both bf16 planes of Wm, up to 128 KB) is built once per CTA from the fp32 parameters.
// The kernels are bound by the traffic of X and Y (128 x (KB + NB) x 256 bytes per tile), not by the MMAs.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_kernels.cuh"
#include "train_kernels.cuh"

namespace tspgnn {

template <int KB, int NB>
struct TRSmem {
  static constexpr int N = 64 * NB;
  static constexpr int W_BYTES = 2 * KB * N * 128;                 // [plane][kb]: image of N output features x 64 k
  static constexpr int SLOT_BYTES = 2 * PLANE_BYTES;               // hi + lo planes of one k-block of a tile
  static constexpr int NSLOT_MAX = (232448 - 1024 - W_BYTES) / SLOT_BYTES;
  static constexpr int NSLOT = NSLOT_MAX > 4 ? 4 : NSLOT_MAX;
  static constexpr int RING_OFF = W_BYTES;
  static constexpr int BAR_OFF = RING_OFF + NSLOT * SLOT_BYTES;
  static constexpr int NBAR = 2 * NSLOT + 4;
  static constexpr int TOTAL = BAR_OFF + 8 * NBAR + 16;
  static constexpr int DYN_BYTES = TOTAL + 128;
  static_assert(NSLOT >= 2, "row GEMM: operand ring needs two slots");
  static_assert(DYN_BYTES <= 232448, "row GEMM: shared memory budget (227 KB)");
};

// 512 threads: the producers are bound by the latency of their loads (Little: 6 TB/s x ~1 us = 40 KB in flight per
// SM), so seven producer warps keep a whole 32 KB k-block of a tile in flight (three warps with eight loads per lane
// held 12 KB: 1.3-2 TB/s).  Registers: 8 x 152 (epilogue) + 8 x 104 (issuer, producers) per lane = the whole file.
constexpr int TR_THREADS = 512;
constexpr int TR_PRODUCERS = 7;

template <int KB, int NB, bool TRANS, int EPI>
__global__ void __launch_bounds__(TR_THREADS, 1) tc_rowgemm_kernel(const RowGemmArgs a) {
  using L = TRSmem<KB, NB>;
  constexpr int N = 64 * NB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* wsm = smem;
  uint8_t* ring = smem + L::RING_OFF;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + L::NSLOT;
  uint64_t* acc_full = empty + L::NSLOT;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles_all = static_cast<int>((a.n_rows + TILE_ROWS - 1) / TILE_ROWS);
  int t0, t1;
  tile_range(blockIdx.x, gridDim.x, n_tiles_all, t0, t1);
  const int ntiles = t1 - t0;

  if (tid == 0) {
    for (int s = 0; s < L::NSLOT; ++s) {
      ptx::mbar_init(&full[s], TR_PRODUCERS);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc(tmem_slot, 512);
  // ---- B operand: Wm[k, n] (TRANS: Ws[n, k], else Ws[k, n]; zero outside w_rows x w_cols), bf16 hi / lo images.
  // One 16-byte chunk = 8 consecutive k of one output feature n; consecutive threads take consecutive n.
  for (int i = tid; i < KB * 8 * N; i += TR_THREADS) {
    const int n = i % N, kc = i / N;              // kc: chunk of 8 k-values, 0 .. 8 KB - 1
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kc * 8 + j;
      float x = 0.f;
      if (!TRANS) {
        if (k < a.w_rows && n < a.w_cols) x = __ldg(a.w + static_cast<int64_t>(k) * a.ldw + n);
      } else {
        if (n < a.w_rows && k < a.w_cols) x = __ldg(a.w + static_cast<int64_t>(n) * a.ldw + k);
      }
      v[j] = x;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const int kb = kc >> 3, kin = kc & 7;
    uint8_t* p = wsm + static_cast<size_t>(kb) * (N * 128) + kin * (N * 16) + n * 16;
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + static_cast<size_t>(KB) * (N * 128)) = lo;
  }
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    ptx::setmaxnreg_inc<152>();
    // ---- epilogue: accumulator -> (+bias, ReLU / mask / accumulate) -> row-major fp32 Y -----------------
    const int e = warp >> 2, q4 = warp & 3;
    const int tq0 = lane & 3, tq1 = lane >> 2;
    const uint32_t t_acc = tmem + (static_cast<uint32_t>(q4 * 32) << 16) + e * 256;
    int use = 0;
    long long* tl = (q4 == 0 && lane == 0) ? a.timeline : nullptr;
    for (int n = e; n < ntiles; n += 2, ++use) {
      tl_mark(tl, e, n, 0);
      const int64_t row0 = static_cast<int64_t>(t0 + n) * TILE_ROWS + q4 * 32 + tq1;   // + 8 m
      // ReLU mask / previous output of a 64-wide layer: requested before the accumulator is waited for
      // (otherwise every tile pays a second memory round trip between the TMEM read and the store)
      constexpr bool PRE = (NB == 1) && ((EPI & (EPI_MASK | EPI_ACCUM)) != 0);
      float2 pre[PRE ? 4 : 1][PRE ? 8 : 1];
      if (PRE) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int64_t row = row0 + 8 * m;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            pre[m][k] = make_float2((EPI & EPI_MASK) ? 1.f : 0.f, (EPI & EPI_MASK) ? 1.f : 0.f);
            if (row < a.n_rows) {
              const float* src = (EPI & EPI_MASK) ? a.mask + row * a.mld : a.y[0] + row * a.yld[0];
              pre[m][k] = *reinterpret_cast<const float2*>(src + 2 * tq0 + 8 * k);
            }
          }
        }
      }
      ptx::mbar_wait(&acc_full[e], use & 1);
      tl_mark(tl, e, n, 1);
      ptx::tcgen05_fence_after();
#pragma unroll 1
      for (int nbi = 0; nbi < NB; ++nbi) {
        // CTAs start at different 64-column blocks: with every CTA on block 0, then 1, ... all accesses of the machine
        // fall into the same quarter of the 1 KB rows of a 256-wide matrix at any one time
        const int nb = (nbi + static_cast<int>(blockIdx.x)) % NB;
        float qa[32], qb[32];     // rows tq1, tq1 + 8 (qa) and tq1 + 16, tq1 + 24 (qb); reg[4k + 2hi + j] = column 8k + 2 tq0 + j
        ptx::tmem_ld_quad64(t_acc + nb * 64, qa, qb);
        if (nbi == NB - 1) {      // last TMEM read of this tile: hand the accumulator back
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_empty[e]);
        }
        float* yp = a.y[nb];
        const int yld = a.yld[nb];
        float2 bq[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int col = nb * 64 + 8 * k + 2 * tq0;
          bq[k].x = (a.bias != nullptr && col < a.bias_n) ? __ldg(a.bias + col) : 0.f;
          bq[k].y = (a.bias != nullptr && col + 1 < a.bias_n) ? __ldg(a.bias + col + 1) : 0.f;
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int64_t row = row0 + 8 * m;
          if (row < a.n_rows) {
            float* yr = yp + row * yld + 2 * tq0;
            const float* mr = (EPI & EPI_MASK) ? a.mask + row * a.mld + 2 * tq0 : nullptr;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int ix = 4 * k + 2 * (m & 1);
              float2 v = make_float2(((m < 2) ? qa : qb)[ix] + bq[k].x, ((m < 2) ? qa : qb)[ix + 1] + bq[k].y);
              if (EPI & EPI_RELU) v = ptx::relu2(v);
              if (EPI & EPI_MASK) {
                const float2 mk = PRE ? pre[m][k] : *reinterpret_cast<const float2*>(mr + 8 * k);
                v.x = (mk.x > 0.f) ? v.x : 0.f;
                v.y = (mk.y > 0.f) ? v.y : 0.f;
              }
              if (EPI & EPI_ACCUM) {
                const float2 o = (PRE && !(EPI & EPI_MASK)) ? pre[m][k] : *reinterpret_cast<const float2*>(yr + 8 * k);
                v.x += o.x;
                v.y += o.y;
              }
              *reinterpret_cast<float2*>(yr + 8 * k) = v;
            }
          }
        }
      }
      tl_mark(tl, e, n, 2);
    }
  } else if (warp == 8) {
    ptx::setmaxnreg_dec<104>();
    // ---- MMA issuer ----------------------------------------------------------------------------------
    constexpr uint32_t IDESC = ptx::umma_idesc_bf16(128, N);
    const uint64_t adesc0 = ptx::umma_desc_k_nosw(ptx::smem_u32(ring), 2048, 128);
    const uint64_t bdesc0 = ptx::umma_desc_k_nosw(ptx::smem_u32(wsm), N * 16, 128);
    for (int n = 0; n < ntiles; ++n) {
      const int acc = n & 1, k_use = n >> 1;
      if (k_use >= 1) ptx::mbar_wait(&acc_empty[acc], (k_use - 1) & 1);
      const uint32_t d_tmem = tmem + acc * 256;
#pragma unroll 1
      for (int kbi = 0; kbi < KB; ++kbi) {
        const int kb = (kbi + static_cast<int>(blockIdx.x)) % KB;      // the producers' order (see there)
        const int seq = n * KB + kbi, slot = seq % L::NSLOT, suse = seq / L::NSLOT;
        if (lane == 0) tl_mark(a.timeline, 2, seq, 0);
        ptx::mbar_wait(&full[slot], suse & 1);
        if (lane == 0) tl_mark(a.timeline, 2, seq, 1);
        ptx::tcgen05_fence_after();
        if (ptx::elect_one()) {
          const int pa_[3] = {1, 0, 0}, pb_[3] = {0, 1, 0};        // (A plane, B plane): lo.hi, hi.lo, hi.hi
          const uint64_t aslot = adesc0 + static_cast<uint32_t>((slot * L::SLOT_BYTES) >> 4);
#pragma unroll
          for (int cb = 0; cb < 3; ++cb) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_ss(d_tmem, aslot + ((pa_[cb] * PLANE_BYTES + k * 4096) >> 4),
                                bdesc0 + static_cast<uint32_t>(((pb_[cb] * KB + kb) * (N * 128) + k * (2 * N * 16)) >> 4), IDESC,
                                (kbi | cb | k) ? 1u : 0u);
          }
          ptx::umma_commit(&empty[slot]);
          if (kbi == KB - 1) ptx::umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (lane == 0) tl_mark(a.timeline, 2, seq, 2);
      }
    }
  } else {
    // ---- producers: X block -> bf16 hi / lo operand planes ------------------------------------------------
    ptx::setmaxnreg_dec<104>();
    const int gw = warp - 9;
    const int r8 = lane & 7, cq = lane >> 3;
    constexpr int GPW = (16 + TR_PRODUCERS - 1) / TR_PRODUCERS;     // 8-row groups per warp (3; 16 groups in a tile)
    for (int n = 0; n < ntiles; ++n) {
      const int64_t row0 = static_cast<int64_t>(t0 + n) * TILE_ROWS;
#pragma unroll 1
      for (int kbi = 0; kbi < KB; ++kbi) {
        const int kb = (kbi + static_cast<int>(blockIdx.x)) % KB;      // CTAs start at different k-blocks (see the epilogue)
        const int seq = n * KB + kbi, slot = seq % L::NSLOT, suse = seq / L::NSLOT;
        const float* xp = a.x[kb];
        const int xld = a.xld[kb];
        const uint32_t slot_s = ptx::smem_u32(ring + slot * L::SLOT_BYTES);
        // every load of this warp's share of the block first (up to twelve 16-byte loads per lane), then the slot
        // wait: the loads do not touch the slot
        long long* tlp = (gw == 0 && lane == 0) ? a.timeline : nullptr;
        tl_mark(tlp, 3, seq, 0);
        float u[GPW][2][8];                             // one 32-byte load (LDG.256) per lane and chunk
#pragma unroll
        for (int i = 0; i < GPW; ++i) {
          const int g = gw + TR_PRODUCERS * i;
          const int64_t row = row0 + g * 8 + r8;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 8; ++e) u[i][j][e] = 0.f;
            if (g < 16 && row < a.n_rows) ptx::ldg256_coherent(xp + row * xld + (cq + 4 * j) * 8, u[i][j]);
          }
        }
        tl_mark(tlp, 3, seq, 1);
        if (suse >= 1) ptx::mbar_wait(&empty[slot], (suse - 1) & 1);
        tl_mark(tlp, 3, seq, 2);
#pragma unroll
        for (int i = 0; i < GPW; ++i) {
          const int g = gw + TR_PRODUCERS * i;
          if (i == 1) tl_mark(tlp, 3, seq, 3);            // the first group's data has arrived and is converted
          if (g < 16) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint4 hi, lo;
              split8(u[i][j], hi, lo);
              const uint32_t off = slot_s + (cq + 4 * j) * 2048 + (g * 8 + r8) * 16;
              ptx::sts128(off, hi);
              ptx::sts128(off + PLANE_BYTES, lo);
            }
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&full[slot]);
        tl_mark(tlp, 3, seq, 4);
      }
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc(tmem, 512);
}


// =====================================================================================================
// dW[kb*64 + k, nb*64 + n] += sum_r X[r, kb*64 + k] * dY[r, nb*64 + n],   db[n] += sum_r dY[r, n]
// on tcgen05 (the contract of xtdy_kernel: every kernel / bias gradient of the reverse pass).
//
// The contraction runs over ROWS.  The chunk-major tile image the forward kernels use as a K-major A operand
// (element (row, col) at (col / 8) * 2048 + row * 16 + (col % 8) * 2) is at the same time the canonical
// un-swizzled MN-major layout of the matrix [col][row]: 8 consecutive rows (k) of 8 columns (mn) form the
// 128-byte core matrix, core matrices are 128 bytes apart along k (LBO) and 2048 bytes apart along mn (SBO).
// So one image of the X tile is the A operand (M = features of X, K = rows) and one image of the dY tile the
// B operand (N = 64 columns of dY, K = rows) of  D[M, N] += A . B^T  with both major bits set in the
// instruction descriptor; a 128-row tile is 8 MMAs of K = 16 per operand-plane combination, bf16x3 as everywhere.
// D accumulates in TMEM over all tiles of the CTA (128 lanes x 64 columns).
//
// Grid (gx, 64-column blocks of dY): CTA (bx, nb) owns a contiguous range of row tiles.  288 threads:
//   warps 0-7 : producers (X and dY tiles: fp32 -> bf16 hi / lo images, two stages of 96 KB); afterwards
//               warps 0-3 read the accumulator (lane = feature) and write the CTA's partial to scratch
//   warp  8   : tcgen05.mma issuer
// The bias gradient is the row of D that belongs to an extra all-ones feature of X (feature 64; only when
// X has 64 features: every biased layer of the model does).  CTA bx adds its partial to slot bx of a blob of
// per-CTA partial gradients laid out like the parameters ([slots][total]; launches of one stream are ordered and
// the two streams of the reverse pass touch different tensors, so plain read-modify-write is enough);
// grad_partial_reduce_kernel sums the slots in fixed order once at the end of the reverse pass: deterministic, no
// floating-point atomics, and no per-launch reduction (a last-CTA reduction of 32 partials cost 30 us per launch).
// =====================================================================================================
constexpr int XT2_THREADS = 288;

struct Xtdy2Args {
  XtdyArgs x;            // the contract of xtdy_kernel (dw / db are only used for their offsets, see below)
  int kblocks;           // 64-column blocks of X (1 or 2)
  float* pblob;          // [slots][total] per-CTA partial gradients, slot = blockIdx.x
  int64_t total;         // floats per slot (= size of the parameter blob)
  int64_t dw_off, db_off;  // offsets of dW / db inside a slot (db_off < 0: no bias)
};

// SROWS rows per stage, NCH 8-column chunks of dY per CTA:
//   <128, 8>  : 64 columns of dY per CTA (grid.y = 64-column blocks), 128-row stages -- the 64-wide MLP layers
//   <64, 32>  : all 256 columns of dY in one CTA (one read of X instead of four, N = 256 MMAs at full rate),
//               64-row stages -- the LSTM kernel gradient dK += [x, h]^T . dz
// Either way a stage is X hi/lo (128 features) + dY hi/lo = 96 KB, two stages.
template <int SROWS, int NCH>
struct XT2Smem {
  static constexpr int CS = SROWS * 16;                  // bytes between 8-column chunks of an image
  static constexpr int XP = 16 * CS;                     // one bf16 plane of the X image (128 features)
  static constexpr int YP = NCH * CS;                    // one bf16 plane of the dY image
  static constexpr int STAGE = 2 * XP + 2 * YP;
  static constexpr int BAR_OFF = 2 * STAGE;
  static constexpr int TOTAL = BAR_OFF + 8 * 5 + 16;
  static constexpr int DYN_BYTES = TOTAL + 128;
  static_assert(DYN_BYTES <= 232448, "xtdy: shared memory budget (227 KB)");
};

// kind::f16 instruction descriptor, bf16 A / B, fp32 D, BOTH operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int m, int n) {
  return ptx::umma_idesc_bf16(m, n) | (1u << 15) | (1u << 16);
}

template <int SROWS, int NCH>
__global__ void __launch_bounds__(XT2_THREADS, 1) tc_xtdy_kernel(const Xtdy2Args q) {
  using L = XT2Smem<SROWS, NCH>;
  constexpr int NCOLS = NCH * 8;
  const XtdyArgs& a = q.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);   // [2]
  uint64_t* empty = full + 2;                                        // [2]
  uint64_t* acc_full = empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int col0 = blockIdx.y * NCOLS;                  // first column of dY / dW of this CTA
  const int KB = q.kblocks;
  const bool ones = (a.db != nullptr) && KB == 1;      // bias gradient through the all-ones feature 64
  const int n_tiles_all = static_cast<int>((a.n_rows + SROWS - 1) / SROWS);
  int t0, t1;
  tile_range(blockIdx.x, gridDim.x, n_tiles_all, t0, t1);
  const int ntiles = t1 - t0;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&full[s], 8);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc(tmem_slot, NCOLS);
  // feature chunks the producers never write (X has 64 or 128 features, M is always 128) must read as zero
  for (int i = tid; i < 2 * L::STAGE / 16; i += XT2_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // ---- producers ------------------------------------------------------------------------------------
    // a warp task = 8 rows x 4 chunks of 8 columns (lane -> row lane & 7, chunk lane >> 3): every lane reads 32
    // contiguous bytes, every quarter-warp writes 128 contiguous bytes of an operand plane
    const int r8 = lane & 7, cq = lane >> 3;
    const int quads = 2 * KB + NCH / 4;                 // chunk quads per row group: X (2 per 64 columns), dY
    const int ntask = (SROWS / 8) * quads;
    const float* dyp = a.dy + col0;
    for (int n = 0; n < ntiles; ++n) {
      const int st = n & 1, use = n >> 1;
      const int64_t row0 = static_cast<int64_t>(t0 + n) * SROWS;
      const uint32_t xs = ptx::smem_u32(smem + st * L::STAGE), ys = xs + 2 * L::XP;
      // every load of the stage first (up to twelve tasks = twenty-four 16-byte loads per lane: the producers are bound
      // by the latency of their loads, two batches per stage made a stage two round trips), then the stage wait
      constexpr int TPW = 12;                           // tasks per warp: <64,32> with two k-blocks has 96, <128,8> 64 or 96
      float u[TPW][8];                                  // one 32-byte load (LDG.256) per task and lane
      uint32_t dst[TPW];
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        const int task = warp + 8 * i;
#pragma unroll
        for (int e = 0; e < 8; ++e) u[i][e] = 0.f;
        dst[i] = 0u;
        if (task < ntask) {
          const int g = task / quads, qd = task % quads;
          const int64_t row = row0 + g * 8 + r8;
          const float* p;
          if (qd < 2 * KB) {
            const int kb = qd >> 1, chunk = (qd & 1) * 4 + cq;           // chunk inside the 64-column block
            p = (kb == 0 ? a.x[0] : a.x[1]) + row * (kb == 0 ? a.xld[0] : a.xld[1]) + chunk * 8;
            dst[i] = xs + (kb * 8 + chunk) * L::CS + (g * 8 + r8) * 16;
          } else {
            const int chunk = (qd - 2 * KB) * 4 + cq;
            p = dyp + row * a.dyld + chunk * 8;
            dst[i] = ys + chunk * L::CS + (g * 8 + r8) * 16;
          }
          if (row < a.n_rows) ptx::ldg256_coherent(p, u[i]);
        }
      }
      if (use >= 1) ptx::mbar_wait(&empty[st], (use - 1) & 1);
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        if (warp + 8 * i < ntask) {
          uint4 hi, lo;
          split8(u[i], hi, lo);
          const bool is_x = dst[i] < ys;
          ptx::sts128(dst[i], hi);
          ptx::sts128(dst[i] + (is_x ? L::XP : L::YP), lo);
        }
      }
      if (ones && warp < SROWS / 32) {
        // feature 64 (chunk 8, element 0) = 1 for the valid rows of the tile: its row of D is colsum(dY)
        const int r = warp * 32 + lane;
        const uint32_t one = (row0 + r < a.n_rows) ? 0x3F80u : 0u;       // bf16(1.0) in the low half
        ptx::sts128(xs + 8 * L::CS + r * 16, make_uint4(one, 0u, 0u, 0u));
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full[st]);
    }
    // ---- partial of this CTA: lane = feature, NCOLS columns, added to this CTA's slot ---------------------------
    if (warp < 4 && ntiles > 0) {
      ptx::mbar_wait(acc_full, 0);
      ptx::tcgen05_fence_after();
      const int m = warp * 32 + lane;
      float* slot = q.pblob + static_cast<int64_t>(blockIdx.x) * q.total;
#pragma unroll 1
      for (int cb = 0; cb < NCOLS / 64; ++cb) {
        float v[64];
        ptx::tmem_ld64(tmem + (static_cast<uint32_t>(warp * 32) << 16) + cb * 64, v);
        const int c0 = col0 + cb * 64;
        float* dstp = nullptr;
        if (m < 64 * KB && m < a.w_rows) dstp = slot + q.dw_off + static_cast<int64_t>(m) * a.ldw + c0;
        else if (ones && m == 64) dstp = slot + q.db_off + c0;
        int ncols = a.w_cols - c0;
        ncols = ncols > 64 ? 64 : ncols;
        if (dstp != nullptr && ncols > 0) {
          if ((ncols & 3) == 0 && (reinterpret_cast<uintptr_t>(dstp) & 15) == 0) {
            // all sixteen loads first: left to the compiler the read-modify-writes were serialised (one memory
            // round trip each, 50 us per launch)
            float4 o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = (4 * j < ncols) ? __ldcg(reinterpret_cast<const float4*>(dstp) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (4 * j < ncols) {
                o[j].x += v[4 * j]; o[j].y += v[4 * j + 1]; o[j].z += v[4 * j + 2]; o[j].w += v[4 * j + 3];
                reinterpret_cast<float4*>(dstp)[j] = o[j];
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (j < ncols) dstp[j] += v[j];
          }
        }
      }
    }
  } else {
    // ---- MMA issuer ----------------------------------------------------------------------------------------
    constexpr uint32_t IDESC = umma_idesc_bf16_mn(128, NCOLS);
    for (int n = 0; n < ntiles; ++n) {
      const int st = n & 1, use = n >> 1;
      ptx::mbar_wait(&full[st], use & 1);
      ptx::tcgen05_fence_after();
      if (ptx::elect_one()) {
        const uint32_t xs = ptx::smem_u32(smem + st * L::STAGE), ys = xs + 2 * L::XP;
        // MN-major: LBO = 128 (next 8 rows), SBO = chunk stride (next 8 columns)
        const uint64_t ad = ptx::umma_desc_k_nosw(xs, 128, L::CS);
        const uint64_t bd = ptx::umma_desc_k_nosw(ys, 128, L::CS);
        const int pa_[3] = {1, 0, 0}, pb_[3] = {0, 1, 0};          // (X plane, dY plane): lo.hi, hi.lo, hi.hi
#pragma unroll
        for (int cb = 0; cb < 3; ++cb) {
#pragma unroll
          for (int k = 0; k < SROWS / 16; ++k)
            ptx::umma_bf16_ss(tmem, ad + ((pa_[cb] * L::XP + k * 256) >> 4), bd + ((pb_[cb] * L::YP + k * 256) >> 4), IDESC,
                              (n | cb | k) ? 1u : 0u);
        }
        ptx::umma_commit(&empty[st]);
        if (n == ntiles - 1) ptx::umma_commit(acc_full);
      }
      __syncwarp();
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) ptx::tmem_dealloc(tmem, NCOLS);
}

// grads[i] += sum over the slots of the per-CTA partial gradients, in slot order
__global__ void __launch_bounds__(256) grad_partial_reduce_kernel(const float* __restrict__ pblob, int slots, int64_t stride,
                                                                  int64_t total, float* __restrict__ grads) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
  int b = 0;
  for (; b + 4 <= slots; b += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) s4[j] += pblob[static_cast<int64_t>(b + j) * stride + i];
  }
  for (; b < slots; ++b) s4[b & 3] += pblob[static_cast<int64_t>(b) * stride + i];
  grads[i] += (s4[0] + s4[1]) + (s4[2] + s4[3]);
}


// =====================================================================================================
// One layer of an MLP's reverse chain in ONE kernel (mlp_reverse):
//   dW_l += a_{l-1}^T . d_l,  db_l += colsum(d_l)          (tc_xtdy_kernel's contract)
//   d_{l-1} = (d_l . W_l^T) * (a_{l-1} > 0)                 (tc_rowgemm_kernel<1,1,TRANS,MASK>; l = 0: plain or +=)
// Both products consume the SAME shared-memory image of the d_l tile: as the K-major A operand of d_l . W_l^T
// (M = rows, K = output features) and as the MN-major B operand of a^T . d_l (N = output features, K = rows);
// a_{l-1} is read once, as the MN-major A operand image and (fp32, from L2) as the ReLU mask.  Against the two
// separate kernels a layer moves 77 MB instead of 128 MB at the north-star size and costs one launch.
// 416 threads: warps 0-3 epilogue (d_{l-1} tile; at the end the CTA's weight-gradient partial), warps 4-11
// producers, warp 12 MMA issuer.  TMEM: two 64-column accumulators for d_l . W^T and one for a^T . d_l.
// =====================================================================================================
constexpr int LR_THREADS = 416;

struct LayerRevArgs {
  const float* d;        // d_l [rows, 64], leading dimension dld
  int dld;
  const float* a;        // a_{l-1} [rows, 64]: input of layer l
  int ald;
  // alternative to `a`: a_{l-1} as bf16 hi / lo tile images written by the training forward (K2Args::act_out),
  // tile t at a_img + t * a_img_stride: [hi 16 KB | lo 16 KB].  The stage takes them by bulk copy (no conversion) and
  // the ReLU mask is read from the hi plane (a > 0 <=> bf16(a) != 0 for a ReLU output).
  const uint8_t* a_img;
  int64_t a_img_stride;
  long long* timeline;   // optional clock64() trace (tools/timeline_train.py), nullptr in production
  float* y;              // d_{l-1} [rows, 64]
  int yld;
  const float* w;        // W_l stored [w_rows = inputs][w_cols = outputs], leading dimension ldw
  int ldw, w_rows, w_cols;
  float* pblob;          // per-CTA partial gradients (see tc_xtdy_kernel)
  int64_t total, dw_off, db_off;
  int64_t n_rows;
  // d_l and d_{l-1} as operand images too (one per 128-row tile: [hi 16 KB | lo 16 KB], tile t at base + t * stride):
  // inside an MLP's reverse chain the layer kernels hand d from one to the next in this form.  The consumer takes it
  // by bulk copy (no producer work); the epilogue writes it with 4-byte stores that cover one full 128-byte line
  // per warp instruction (8 rows x 16 bytes of a chunk), where the row-major fp32 form touches 8 lines for 256
  // bytes -- the epilogue's load/store wavefronts bounded the kernel (tools/timeline_train.py 7).  Rows past n_rows
  // are written as zeros.
  const uint8_t* d_img;
  int64_t d_img_stride;
  uint8_t* y_img;
  int64_t y_img_stride;
};

struct LRSmem {
  static constexpr int W_BYTES = 2 * 64 * 128;                       // W^T image, hi + lo
  static constexpr int DP = PLANE_BYTES;                             // one plane of the d image (64 columns)
  static constexpr int AP = 2 * PLANE_BYTES;                         // one plane of the a image (128 features: 64 + ones + zeros)
  static constexpr int STAGE = 2 * DP + 2 * AP;                      // 96 KB
  static constexpr int STAGE_OFF = W_BYTES;
  static constexpr int BAR_OFF = STAGE_OFF + 2 * STAGE;
  static constexpr int TOTAL = BAR_OFF + 8 * 9 + 16;
  static constexpr int DYN_BYTES = TOTAL + 128;
};
static_assert(LRSmem::DYN_BYTES <= 232448, "layer reverse: shared memory budget (227 KB)");

template <int EPI>
__global__ void __launch_bounds__(LR_THREADS, 1) tc_layer_reverse_kernel(const LayerRevArgs a) {
  using L = LRSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* wsm = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);   // [2] stage filled (8 producer warps)
  uint64_t* empty = full + 2;                                        // [2] stage consumed (MMA commit)
  uint64_t* d1_full = empty + 2;                                     // [2] d_l . W^T of a tile complete
  uint64_t* d1_empty = d1_full + 2;                                  // [2] its accumulator drained (4 epilogue warps)
  uint64_t* d2_full = d1_empty + 2;                                  // a^T . d_l of all tiles complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ones = a.db_off >= 0;
  const int n_tiles_all = static_cast<int>((a.n_rows + TILE_ROWS - 1) / TILE_ROWS);
  int t0, t1;
  tile_range(blockIdx.x, gridDim.x, n_tiles_all, t0, t1);
  const int ntiles = t1 - t0;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&full[s], 8);
      ptx::mbar_init(&empty[s], 1);
      ptx::mbar_init(&d1_full[s], 1);
      ptx::mbar_init(&d1_empty[s], 4);
    }
    ptx::mbar_init(d2_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 12) ptx::tmem_alloc(tmem_slot, 256);
  // the a-image chunks the producers never write (features 72..127) must read as zero
  for (int i = tid; i < 2 * L::STAGE / 16; i += LR_THREADS)
    reinterpret_cast<uint4*>(smem + L::STAGE_OFF)[i] = make_uint4(0u, 0u, 0u, 0u);
  // B operand of d_l . W^T: image[n = input i][k = output o] = W[i][o]
  for (int i = tid; i < 8 * 64; i += LR_THREADS) {
    const int n = i & 63, kc = i >> 6;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kc * 8 + j;
      v[j] = (n < a.w_rows && k < a.w_cols) ? __ldg(a.w + static_cast<int64_t>(n) * a.ldw + k) : 0.f;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    uint8_t* p = wsm + kc * (64 * 16) + n * 16;
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + 64 * 128) = lo;
  }
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ---- epilogue: d_{l-1} tile ---------------------------------------------------------------------------------
    const int tq0 = lane & 3, tq1 = lane >> 2;
    long long* tl = (warp == 0 && lane == 0) ? a.timeline : nullptr;
    for (int n = 0; n < ntiles; ++n) {
      tl_mark(tl, 0, n, 0);
      const int acc = n & 1, use = n >> 1;
      const int64_t row0 = static_cast<int64_t>(t0 + n) * TILE_ROWS + warp * 32 + tq1;   // + 8 m
      float2 pre[4][8];           // ReLU mask source a_{l-1} (MASK) or the previous output (ACCUM), before the wait
      if ((EPI & EPI_MASK) && a.a_img != nullptr) {
        // the mask from the hi plane of the image: chunk k, row, columns 2 tq0, 2 tq0 + 1 = one 32-bit word
        const uint8_t* img = a.a_img + static_cast<int64_t>(t0 + n) * a.a_img_stride;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int rt = warp * 32 + tq1 + 8 * m;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t w2 = __ldg(reinterpret_cast<const uint32_t*>(img + k * 2048 + rt * 16 + tq0 * 4));
            pre[m][k] = make_float2((w2 & 0xffffu) ? 1.f : 0.f, (w2 >> 16) ? 1.f : 0.f);
          }
        }
      } else if (EPI & (EPI_MASK | EPI_ACCUM)) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int64_t row = row0 + 8 * m;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            pre[m][k] = make_float2(0.f, 0.f);
            if (row < a.n_rows) {
              const float* src = (EPI & EPI_MASK) ? a.a + row * a.ald : a.y + row * a.yld;
              pre[m][k] = *reinterpret_cast<const float2*>(src + 2 * tq0 + 8 * k);
            }
          }
        }
      }
      tl_mark(tl, 0, n, 1);
      ptx::mbar_wait(&d1_full[acc], use & 1);
      tl_mark(tl, 0, n, 2);
      ptx::tcgen05_fence_after();
      float qa[32], qb[32];
      ptx::tmem_ld_quad64(tmem + (static_cast<uint32_t>(warp * 32) << 16) + acc * 64, qa, qb);
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&d1_empty[acc]);
      if (a.y_img != nullptr) {
        // d_{l-1} as hi / lo image: word (chunk k, row, column pair tq0); a warp store covers 128 contiguous bytes
        uint8_t* img = a.y_img + static_cast<int64_t>(t0 + n) * a.y_img_stride;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int rt = warp * 32 + tq1 + 8 * m;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int ix = 4 * k + 2 * (m & 1);
            float2 v = make_float2(((m < 2) ? qa : qb)[ix], ((m < 2) ? qa : qb)[ix + 1]);
            if (EPI & EPI_MASK) {
              v.x = (pre[m][k].x > 0.f) ? v.x : 0.f;
              v.y = (pre[m][k].y > 0.f) ? v.y : 0.f;
            }
            uint32_t hi, lo;
            ptx::split_bf16x2(v.x, v.y, hi, lo);
            uint32_t* w = reinterpret_cast<uint32_t*>(img + k * 2048 + rt * 16 + tq0 * 4);
            w[0] = hi;
            w[PLANE_BYTES / 4] = lo;
          }
        }
        tl_mark(tl, 0, n, 3);
        continue;
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int64_t row = row0 + 8 * m;
        if (row < a.n_rows) {
          float* yr = a.y + row * a.yld + 2 * tq0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int ix = 4 * k + 2 * (m & 1);
            float2 v = make_float2(((m < 2) ? qa : qb)[ix], ((m < 2) ? qa : qb)[ix + 1]);
            if (EPI & EPI_MASK) {
              v.x = (pre[m][k].x > 0.f) ? v.x : 0.f;
              v.y = (pre[m][k].y > 0.f) ? v.y : 0.f;
            }
            if (EPI & EPI_ACCUM) {
              v.x += pre[m][k].x;
              v.y += pre[m][k].y;
            }
            *reinterpret_cast<float2*>(yr + 8 * k) = v;
          }
        }
      }
      tl_mark(tl, 0, n, 3);
    }
    // ---- the CTA's weight-gradient partial: lane = input feature, 64 output columns ---------------------------------
    if (ntiles > 0) {
      ptx::mbar_wait(d2_full, 0);
      ptx::tcgen05_fence_after();
      float v[64];
      ptx::tmem_ld64(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 128, v);
      const int m = warp * 32 + lane;
      float* slot = a.pblob + static_cast<int64_t>(blockIdx.x) * a.total;
      float* dstp = nullptr;
      if (m < 64 && m < a.w_rows) dstp = slot + a.dw_off + static_cast<int64_t>(m) * a.ldw;
      else if (ones && m == 64) dstp = slot + a.db_off;
      const int ncols = a.w_cols > 64 ? 64 : a.w_cols;
      if (dstp != nullptr) {
        if ((ncols & 3) == 0 && (reinterpret_cast<uintptr_t>(dstp) & 15) == 0) {
          float4 o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = (4 * j < ncols) ? __ldcg(reinterpret_cast<const float4*>(dstp) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (4 * j < ncols) {
              o[j].x += v[4 * j]; o[j].y += v[4 * j + 1]; o[j].z += v[4 * j + 2]; o[j].w += v[4 * j + 3];
              reinterpret_cast<float4*>(dstp)[j] = o[j];
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j)
            if (j < ncols) dstp[j] += v[j];
        }
      }
    }
  } else if (warp < 12) {
    // ---- producers: d_l and a_{l-1} tiles -> bf16 hi / lo images ----------------------------------------------------
    const int pw = warp - 4;
    const int r8 = lane & 7, cq = lane >> 3;
    constexpr int quads = 4, ntask = 16 * quads;          // per 8-row group: d (2 quads of 4 chunks), a (2)
    for (int n = 0; n < ntiles; ++n) {
      const int st = n & 1, use = n >> 1;
      long long* tlp = (pw == 0 && lane == 0) ? a.timeline : nullptr;
      tl_mark(tlp, 3, n, 0);
      const int64_t row0 = static_cast<int64_t>(t0 + n) * TILE_ROWS;
      const uint32_t ds = ptx::smem_u32(smem + L::STAGE_OFF + st * L::STAGE), as = ds + 2 * L::DP;
      // the tile's sixteen 16-byte loads per lane are requested before the stage is waited for
      // with a_img only the d tasks are left (two quads per row group): the a image comes by bulk copy
      // operands that come as images are bulk-copied below: only the row-major ones are left as tasks
      static_assert(quads == 4, "d tasks are quads 0-1, a tasks quads 2-3");
      const int qoff = (a.d_img != nullptr) ? 2 : 0;
      const int quads_eff = quads - ((a.d_img != nullptr) ? 2 : 0) - ((a.a_img != nullptr) ? 2 : 0);
      float u[8][8];                                    // one 32-byte load (LDG.256) per task and lane
      uint32_t dst[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int task = pw + 8 * i;
        if (task >= 16 * quads_eff) {
          dst[i] = 0u;
          continue;
        }
        const int g = task / quads_eff, qd = task % quads_eff + qoff;
        const int64_t row = row0 + g * 8 + r8;
        const int chunk = (qd & 1) * 4 + cq;
        const float* p = (qd < 2) ? a.d + row * a.dld + chunk * 8 : a.a + row * a.ald + chunk * 8;
        dst[i] = ((qd < 2) ? ds : as) + chunk * 2048 + (g * 8 + r8) * 16;
#pragma unroll
        for (int e = 0; e < 8; ++e) u[i][e] = 0.f;
        if (row < a.n_rows) ptx::ldg256_coherent(p, u[i]);
      }
      static_assert(ntask == 64, "eight producer warps, eight tasks each");
      tl_mark(tlp, 3, n, 1);
      if (use >= 1) ptx::mbar_wait(&empty[st], (use - 1) & 1);
      tl_mark(tlp, 3, n, 2);
      if (pw == 0 && lane == 0) {
        uint8_t* stage = smem + L::STAGE_OFF + st * L::STAGE;
        if (a.a_img != nullptr) {
          // hi plane -> chunks 0..7 of the a image's hi plane, lo plane likewise (chunks 8..15 stay ones / zeros)
          const uint8_t* img = a.a_img + static_cast<int64_t>(t0 + n) * a.a_img_stride;
          ptx::mbar_expect_tx(&full[st], 2 * PLANE_BYTES);
          ptx::bulk_g2s(stage + 2 * L::DP, img, PLANE_BYTES, &full[st]);
          ptx::bulk_g2s(stage + 2 * L::DP + L::AP, img + PLANE_BYTES, PLANE_BYTES, &full[st]);
        }
        if (a.d_img != nullptr) {
          const uint8_t* img = a.d_img + static_cast<int64_t>(t0 + n) * a.d_img_stride;
          ptx::mbar_expect_tx(&full[st], 2 * PLANE_BYTES);
          ptx::bulk_g2s(stage, img, 2 * PLANE_BYTES, &full[st]);       // hi | lo = the d image of the stage (DP apart)
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (dst[i] == 0u) continue;
        uint4 hi, lo;
        split8(u[i], hi, lo);
        const bool is_d = dst[i] < as;
        ptx::sts128(dst[i], hi);
        ptx::sts128(dst[i] + (is_d ? L::DP : L::AP), lo);
      }
      if (ones && pw < 4) {       // feature 64 of the a image = 1 for the valid rows: its row of a^T . d_l is colsum(d_l)
        const int r = pw * 32 + lane;
        const uint32_t one = (row0 + r < a.n_rows) ? 0x3F80u : 0u;
        ptx::sts128(as + 8 * 2048 + r * 16, make_uint4(one, 0u, 0u, 0u));
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full[st]);
      tl_mark(tlp, 3, n, 4);
    }
  } else {
    // ---- MMA issuer -----------------------------------------------------------------------------------------------
    constexpr uint32_t IDESC1 = ptx::umma_idesc_bf16(128, 64);        // d_l . W^T: both K-major
    constexpr uint32_t IDESC2 = umma_idesc_bf16_mn(128, 64);          // a^T . d_l: both MN-major
    const uint64_t wdesc = ptx::umma_desc_k_nosw(ptx::smem_u32(wsm), 64 * 16, 128);
    for (int n = 0; n < ntiles; ++n) {
      const int st = n & 1, use = n >> 1;
      if (lane == 0) tl_mark(a.timeline, 2, n, 0);
      ptx::mbar_wait(&full[st], use & 1);
      if (lane == 0) tl_mark(a.timeline, 2, n, 1);
      if (use >= 1) ptx::mbar_wait(&d1_empty[st], (use - 1) & 1);
      if (lane == 0) tl_mark(a.timeline, 2, n, 2);
      ptx::tcgen05_fence_after();
      if (ptx::elect_one()) {
        const uint32_t ds = ptx::smem_u32(smem + L::STAGE_OFF + st * L::STAGE), as = ds + 2 * L::DP;
        const uint64_t dk = ptx::umma_desc_k_nosw(ds, 2048, 128);     // d image as K-major A (rows x outputs)
        const uint64_t dmn = ptx::umma_desc_k_nosw(ds, 128, 2048);    // d image as MN-major B (outputs x rows)
        const uint64_t amn = ptx::umma_desc_k_nosw(as, 128, 2048);    // a image as MN-major A (inputs x rows)
        const int pa_[3] = {1, 0, 0}, pb_[3] = {0, 1, 0};
#pragma unroll
        for (int cb = 0; cb < 3; ++cb) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss(tmem + st * 64, dk + ((pa_[cb] * L::DP + k * 4096) >> 4),
                              wdesc + ((pb_[cb] * (64 * 128) + k * (2 * 64 * 16)) >> 4), IDESC1, (cb | k) ? 1u : 0u);
        }
        ptx::umma_commit(&d1_full[st]);
#pragma unroll
        for (int cb = 0; cb < 3; ++cb) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            ptx::umma_bf16_ss(tmem + 128, amn + ((pa_[cb] * L::AP + k * 256) >> 4), dmn + ((pb_[cb] * L::DP + k * 256) >> 4),
                              IDESC2, (n | cb | k) ? 1u : 0u);
        }
        ptx::umma_commit(&empty[st]);
        if (n == ntiles - 1) ptx::umma_commit(d2_full);
      }
      __syncwarp();
      if (lane == 0) tl_mark(a.timeline, 2, n, 3);
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 12) ptx::tmem_dealloc(tmem, 256);
}

}  // namespace tspgnn
