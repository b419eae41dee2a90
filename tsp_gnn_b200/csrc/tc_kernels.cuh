// Tensor-core (tcgen05) kernels of the message-passing timestep.
//
//   K1  tc_lnlstm_kernel : gather/segment input -> z = [x,h].K on tcgen05 (TMEM accumulator,
//                          256 columns, double buffered) -> 5x LayerNorm + gates -> new (c,h)
//                          (graphnn.py:155-170 for both variables of model.py:74-92)
//   K2  tc_mlp_kernel    : 4-layer message MLP chained through TMEM (graphnn.py:152-154), then
//                          E rows: scatter-add into xV (= EV^T . msg, model.py:76-83)
//                          V rows: store the vertex message consumed by K1's gather
//                          vote  : E_vote MLP, 64->1 tail in registers (model.py:107-128)
//
// Both are persistent, warp-specialised CTAs of 384 threads (one per SM):
//   warps 0-3, 4-7 : two epilogue warpgroups, one 128-row tile each (thread = row = TMEM lane)
//   warp  8        : tcgen05.mma issuer (one lane), owns the TMEM allocation
//   warps 9-11     : producers (K1: gather of the x operand + bulk copies of h; K2: bulk copies)
//
// HBM layout ("tile images"): the recurrent state of 128 consecutive rows is stored as
//   [h hi : 8 chunks x 128 rows x 16 B (8 bf16)]   16 KB   chunk-major = un-swizzled K-major UMMA
//   [h lo : same, bf16(h - hi)]  (HP == 2 only)    16 KB   canonical layout (LBO 2048, SBO 128)
//   [c    : 16 chunks x 128 rows x 16 B (4 fp32)]  32 KB
// so that (a) one bulk async copy (UBLKCP) stages the h planes as the MMA A operand and (b) the
// epilogue threads (one row each) read c and write c', h' straight from / to global memory with
// fully coalesced 16-byte accesses (a warp covers 512 contiguous bytes per instruction).
// HP = number of bf16 planes: 2 -> every product is hi*hi + hi*lo + lo*hi (fp32-parity mode),
// 1 -> single bf16.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"
#ifndef TSPGNN_DBG_NO_RED
#define TSPGNN_DBG_NO_RED 0
#endif

namespace tspgnn {

constexpr int PLANE_BYTES = TILE_ROWS * 128;   // one bf16 plane of a tile
constexpr int CT_BYTES = TILE_ROWS * 256;      // fp32 c tile
constexpr int TC_THREADS = 384;
constexpr int NUM_GATHER_WARPS = 3;

__host__ __device__ constexpr int tile_bytes(int hp) { return hp * PLANE_BYTES + CT_BYTES; }

// byte offset of element (row, col) inside a bf16 plane / the fp32 c tile
__host__ __device__ inline uint32_t plane_off(int r, int col) {
  return static_cast<uint32_t>((col >> 3) * 2048 + r * 16 + (col & 7) * 2);
}
__host__ __device__ inline uint32_t ct_off(int r, int col) {
  return static_cast<uint32_t>((col >> 2) * 2048 + r * 16 + (col & 3) * 4);
}
// B operand image of n_rows output features x 64 k-values (bf16), same chunk-major layout
__host__ __device__ inline uint32_t wimg_off(int n, int k, int n_rows) {
  return static_cast<uint32_t>((k >> 3) * (n_rows * 16) + n * 16 + (k & 7) * 2);
}

struct K1Args {
  uint8_t* stateE;
  uint8_t* stateV;
  const uint8_t* wE;   // LSTM weight images, [plane][kblock] x (256 x 64 bf16) = 32 KB each
  const uint8_t* wV;
  const float* mV;     // vertex messages [sumV][64] (E rows gather two of them)
  float* xV;           // summed edge messages [sumV_pad][64]; V rows read and then zero it
  const int32_t* src;  // [sumE_pad], zero padded
  const int32_t* dst;
  int64_t nE, nV;
  int tilesE, tilesV, e_ctas;
  int clampE, clampV;    // 1: clamp the logistic exponents of the i / f gates (large LayerNorm gamma / beta)
  // Folded E_msg_V output layer (see K2Args::fold): xV then holds sum_e a3[e] (the last HIDDEN activations
  // of the incident edges), wV is the image of [W4.Kx ; Kh] and the V epilogue adds deg(v) * (b4.Kx),
  // c_vfold_bias, to z before the gate LayerNorms.  nullptr = unfolded.
  const float* vdeg;     // [sumV_pad] number of incident edges of every vertex row
  const float* ln_tab;   // [2 cells][gamma'[5][64] | beta'[5][64]]: LayerNorm parameters as the epilogue wants them (see the prologue)
  long long* timeline;   // optional clock64() trace (tools/timeline.py), nullptr in production
  int tl_slot;           // >= 0: also record %globaltimer at entry / prologue end / dependency wait / exit (launch-gap analysis)
};

struct K2Args {
  const uint8_t* stateE;
  const uint8_t* stateV;
  const uint8_t* wE;   // MLP weight images, [layer][plane] x (64 x 64 bf16) = 8 KB each
  const uint8_t* wV;
  float* xV;
  float* mV;
  float* vote;         // [sumE_pad] (vote mode)
  const int32_t* src;
  const int32_t* dst;
  int64_t nE, nV;
  int tilesE, tilesV, e_ctas;
  int vote_mode;
  // fold = 1: the last (linear) layer of E_msg_V is not applied per edge.  Since
  //   sum_{e in inc(v)} (a3[e].W4 + b4) = (sum_e a3[e]).W4 + deg(v).b4          (graphnn.py:152-161)
  // edge tiles run three layers, scatter the hidden activations a3, and the vertex cell applies
  // W4 (merged into its LSTM kernel) once per vertex instead of once per edge (K1Args::vdeg).
  int fold;
  const float* bias_tab;   // [3 MLPs][4][64] biases (V_msg_E, E_msg_V, E_vote), coalesced copy for the prologue
  // scatter plan (tc_scatter_plan_kernel): per edge tile the 256 (row, endpoint vertex) pairs sorted by vertex
  const uint8_t* ent_row;  // [tilesE][256] row of the tile
  const int32_t* ent_v;    // [tilesE][256] vertex it adds to, -1 = padding
  unsigned int* zero_word; // optional: cleared by this launch (grid barrier counter of the persistent kernel that follows)
  // Training forward (optional): the hidden activations of the EDGE chain (layers 1-3 of E_msg_V, post-ReLU) as bf16
  // hi / lo tile images, [edge tile][layer 0..2][hi 16 KB | lo 16 KB] in the layout of the h planes -- exactly the
  // operand image tc_layer_reverse_kernel needs for a_{l-1}, so the reverse pass neither recomputes nor converts them
  uint8_t* act_out;
  long long* timeline;
  int tl_slot;
};

// (b4 of E_msg_V) . Kx of the V cell, centred per gate like the weight image (K1Args::vdeg)
__constant__ float c_vfold_bias[4 * D];


// timeline slot: [cta][role 0..3][local tile 0..63][event 0..7]
constexpr int TL_ROLES = 4, TL_TILES = 64, TL_EVENTS = 8;
__device__ __forceinline__ void tl_mark(long long* tl, int role, int tile, int ev) {
  if (tl != nullptr && tile < TL_TILES)
    tl[((static_cast<int64_t>(blockIdx.x) * TL_ROLES + role) * TL_TILES + tile) * TL_EVENTS + ev] = clock64();
}

// launch-gap trace: %globaltimer (ns, common to all SMs) of CTA-level events, slot [cta][role 3][tl_slot][ev]
__device__ __forceinline__ void tl_gmark(long long* tl, int slot, int ev) {
  if (tl != nullptr && slot >= 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    tl[((static_cast<int64_t>(blockIdx.x) * TL_ROLES + 3) * TL_TILES + slot) * TL_EVENTS + ev] = static_cast<long long>(t);
  }
}

// contiguous, balanced range of tiles for CTA `i` of `n`
__device__ __forceinline__ void tile_range(int i, int n, int tiles, int& t0, int& t1) {
  t0 = static_cast<int>((static_cast<int64_t>(i) * tiles) / n);
  t1 = static_cast<int>((static_cast<int64_t>(i + 1) * tiles) / n);
}

// 8 fp32 -> one 16-byte chunk of bf16 hi (+ one of bf16 lo)
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  ptx::split_bf16x2(x[0], x[1], hi.x, lo.x);
  ptx::split_bf16x2(x[2], x[3], hi.y, lo.y);
  ptx::split_bf16x2(x[4], x[5], hi.z, lo.z);
  ptx::split_bf16x2(x[6], x[7], hi.w, lo.w);
}

// ====================================================================================
// K1: LayerNorm-LSTM step
// ====================================================================================
template <int HP>
struct K1Smem {
  static constexpr int W_BYTES = HP * 2 * 256 * 128;       // [plane][kblock] 32 KB images
  static constexpr int SLOT_BYTES = HP * PLANE_BYTES;      // all planes of one A k-block (x or h)
  static constexpr int NSLOT = (HP == 2) ? 3 : 6;          // ring of operand slots: x(t), h(t), x(t+1), ...
  static constexpr int RING_OFF = W_BYTES;
  static constexpr int BAR_OFF = RING_OFF + NSLOT * SLOT_BYTES;
  static constexpr int NBAR = 1 + 2 * NSLOT + 4 + 1;
  static constexpr int LN_OFF = (BAR_OFF + 8 * NBAR + 16 + 15) & ~15;   // gamma[5][64], beta[5][64] of this CTA's cell
  static constexpr int TOTAL = LN_OFF + 2 * 5 * D * 4;
  static constexpr int DYN_BYTES = TOTAL + 128;            // slack for 128-B alignment
};
static_assert(K1Smem<2>::DYN_BYTES <= 232448, "K1 shared memory budget (227 KB)");

// ---- producer: build the x operand of one tile in shared memory ------------------------
// E rows: x = mV[src] + mV[dst] (= EV . msg, model.py:85-91); V rows: x = xV (then cleared).
// A warp instruction covers 8 rows x 4 chunks: lane -> (row r8 = lane & 7, chunk cq = lane >> 3),
// so every quarter-warp writes 128 contiguous bytes of shared memory (no bank conflicts).
// NC = true: the messages are read through the non-coherent path (legal when a previous LAUNCH wrote them);
// false: coherent loads, for the persistent kernel whose messages are written during the same launch.
// The gather is bound by L2 latency, not bandwidth (tools/microbench/datapipe.cu: ~800 cycles per round
// trip, 14 B/clk per SM with 4 x 32-byte loads in flight per lane of three warps), so E rows are handled in
// three batches of two 8-row groups: the 8 loads of a batch (64 registers) are all issued before the first
// one is consumed.  k1_gather_load<B> issues batch B, k1_gather_store<B> adds, splits and stores it; the
// caller issues batch 0 BEFORE it waits for the operand slot, so one of the three round trips of a tile
// hides under that wait.
template <bool NC>
__device__ __forceinline__ void k1_gather_load(int b, int gw, int lane, const float* mV, const int (&si)[6],
                                               const int (&di)[6], float (&u)[2][2][8], float (&w)[2][2][8]) {
  const int cq = lane >> 3;
#pragma unroll
  for (int g2 = 0; g2 < 2; ++g2) {
    const int gi = 2 * b + g2;
    if (gw + NUM_GATHER_WARPS * gi < 16) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int chunk = cq + 4 * j;
        if (NC) {
          ptx::ldg256(mV + static_cast<int64_t>(si[gi]) * D + chunk * 8, u[g2][j]);
          ptx::ldg256(mV + static_cast<int64_t>(di[gi]) * D + chunk * 8, w[g2][j]);
        } else {
          ptx::ldg256_coherent(mV + static_cast<int64_t>(si[gi]) * D + chunk * 8, u[g2][j]);
          ptx::ldg256_coherent(mV + static_cast<int64_t>(di[gi]) * D + chunk * 8, w[g2][j]);
        }
      }
    }
  }
}
template <int HP>
__device__ __forceinline__ void k1_gather_store(int b, uint8_t* slot, int gw, int lane, const float (&u)[2][2][8],
                                                const float (&w)[2][2][8]) {
  const int r8 = lane & 7, cq = lane >> 3;
  const uint32_t slot_s = ptx::smem_u32(slot);
#pragma unroll
  for (int g2 = 0; g2 < 2; ++g2) {
    const int g = gw + NUM_GATHER_WARPS * (2 * b + g2);
    if (g < 16) {
      const int row = g * 8 + r8;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int chunk = cq + 4 * j;
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = u[g2][j][i] + w[g2][j][i];
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = slot_s + chunk * 2048 + row * 16;
        ptx::sts128(off, hi);
        if (HP == 2) ptx::sts128(off + PLANE_BYTES, lo);
      }
    }
  }
}

// V rows: x = xV (then cleared).  Same lane -> (row, chunk) mapping; 5 % of the rows, kept simple.
template <int HP>
__device__ __forceinline__ void k1_fill_x_v(uint8_t* slot, int gw, int lane, int64_t row0, float* xV) {
  const int r8 = lane & 7, cq = lane >> 3;
  const uint32_t slot_s = ptx::smem_u32(slot);
#pragma unroll
  for (int gi = 0; gi < 6; ++gi) {
    const int g = gw + NUM_GATHER_WARPS * gi;
    if (g < 16) {
      const int row = g * 8 + r8;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int chunk = cq + 4 * j;
        float x[8];
        float4* p = reinterpret_cast<float4*>(xV + (row0 + row) * D + chunk * 8);
        const float4 u0 = p[0], u1 = p[1];
        p[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        p[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        x[0] = u0.x; x[1] = u0.y; x[2] = u0.z; x[3] = u0.w;
        x[4] = u1.x; x[5] = u1.y; x[6] = u1.z; x[7] = u1.w;
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = slot_s + chunk * 2048 + row * 16;
        ptx::sts128(off, hi);
        if (HP == 2) ptx::sts128(off + PLANE_BYTES, lo);
      }
    }
  }
}

// whole-tile form (all three batches back to back) for callers without a wait to hide a round trip under
template <int HP, bool IS_V, bool NC = true>
__device__ __forceinline__ void k1_fill_x(uint8_t* slot, int gw, int lane, int64_t row0, const float* mV,
                                          float* xV, const int (&si)[6], const int (&di)[6]) {
  if (IS_V) {
    k1_fill_x_v<HP>(slot, gw, lane, row0, xV);
  } else {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      float u[2][2][8], w[2][2][8];
      k1_gather_load<NC>(b, gw, lane, mV, si, di, u, w);
      k1_gather_store<HP>(b, slot, gw, lane, u, w);
    }
  }
}

// The x operand of a CTA's FIRST tile is built by the eight epilogue warps, which have nothing to do
// until the first accumulator is ready and have the registers to keep every load in flight: the
// first gather drops from three L2 round trips to one and the gather warps start on tile 1 at once.
// NW = number of warps that share the tile (warp = 0 .. NW-1): 8 (two 8-row groups each) or 4 (four groups each).
// DEPWAIT: griddepcontrol.wait is executed after the column indices are loaded (they belong to the plan, not
// to the preceding kernel's output) and before the first message is read.
template <int HP, bool IS_V, bool NC = true, int NW = 8, bool DEPWAIT = false>
__device__ __forceinline__ void k1_boot_fill_tile(const float* mV, float* xV,
                                                  const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                  uint8_t* slot, int warp, int lane, int tile) {
  constexpr int G = 16 / NW;
  static_assert(G % 2 == 0, "groups are handled in pairs");
  const int r8 = lane & 7, cq = lane >> 3;
  const uint32_t slot_s = ptx::smem_u32(slot);
  const int64_t row0 = static_cast<int64_t>(tile) * TILE_ROWS;
  int s_[G], d_[G];
#pragma unroll
  for (int gi = 0; gi < G; ++gi) s_[gi] = d_[gi] = 0;
  if (!IS_V) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      s_[gi] = __ldg(src + row0 + (G * warp + gi) * 8 + r8);
      d_[gi] = __ldg(dst + row0 + (G * warp + gi) * 8 + r8);
    }
  }
  if (DEPWAIT) ptx::grid_dependency_wait();
#pragma unroll
  for (int gb = 0; gb < G; gb += 2) {       // two 8-row groups = 8 loads of 32 bytes in flight per lane
    float u[2][2][8], w[2][2][8];
#pragma unroll
    for (int g2 = 0; g2 < 2; ++g2) {
      const int row = (G * warp + gb + g2) * 8 + r8;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int chunk = cq + 4 * j;
        if (IS_V) {
          float4* p = reinterpret_cast<float4*>(xV + (row0 + row) * D + chunk * 8);
          const float4 u0 = p[0], u1 = p[1];
          p[0] = make_float4(0.f, 0.f, 0.f, 0.f);
          p[1] = make_float4(0.f, 0.f, 0.f, 0.f);
          u[g2][j][0] = u0.x; u[g2][j][1] = u0.y; u[g2][j][2] = u0.z; u[g2][j][3] = u0.w;
          u[g2][j][4] = u1.x; u[g2][j][5] = u1.y; u[g2][j][6] = u1.z; u[g2][j][7] = u1.w;
#pragma unroll
          for (int i = 0; i < 8; ++i) w[g2][j][i] = 0.f;
        } else if (NC) {
          ptx::ldg256(mV + static_cast<int64_t>(s_[gb + g2]) * D + chunk * 8, u[g2][j]);
          ptx::ldg256(mV + static_cast<int64_t>(d_[gb + g2]) * D + chunk * 8, w[g2][j]);
        } else {
          ptx::ldg256_coherent(mV + static_cast<int64_t>(s_[gb + g2]) * D + chunk * 8, u[g2][j]);
          ptx::ldg256_coherent(mV + static_cast<int64_t>(d_[gb + g2]) * D + chunk * 8, w[g2][j]);
        }
      }
    }
#pragma unroll
    for (int g2 = 0; g2 < 2; ++g2) {
      const int row = (G * warp + gb + g2) * 8 + r8;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int chunk = cq + 4 * j;
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = u[g2][j][i] + w[g2][j][i];
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = slot_s + chunk * 2048 + row * 16;
        ptx::sts128(off, hi);
        if (HP == 2) ptx::sts128(off + PLANE_BYTES, lo);
      }
    }
  }
}

template <int HP, bool IS_V>
__device__ __forceinline__ void k1_boot_fill(const K1Args& a, uint8_t* slot, int warp, int lane, int tile) {
  k1_boot_fill_tile<HP, IS_V, true, 8, true>(a.mV, a.xV, a.src, a.dst, slot, warp, lane, tile);
}

template <int HP, bool IS_V>
__device__ __forceinline__ void k1_producer(const K1Args& a, uint8_t* state, uint8_t* ring, uint64_t* full,
                                            uint64_t* empty, uint64_t* boot, int t0, int ntiles, int gw, int lane) {
  using L = K1Smem<HP>;
  long long* tl = (gw == 0 && lane == 0) ? a.timeline : nullptr;
  const int r8 = lane & 7;
  int si[6], di[6];
#pragma unroll
  for (int gi = 0; gi < 6; ++gi) si[gi] = di[gi] = 0;
  auto load_idx = [&](int tile, int (&s)[6], int (&d)[6]) {
    if (IS_V) return;
    const int64_t row0 = static_cast<int64_t>(tile) * TILE_ROWS;
#pragma unroll
    for (int gi = 0; gi < 6; ++gi) {
      const int g = gw + NUM_GATHER_WARPS * gi;
      if (g < 16) {
        s[gi] = __ldg(a.src + row0 + g * 8 + r8);
        d[gi] = __ldg(a.dst + row0 + g * 8 + r8);
      }
    }
  };
  for (int n = 0; n < ntiles; ++n) {   // (tile 0 needs no indices here: the epilogue warps build its operand)
    const int tile = t0 + n;
    // ---- h operand (k-block 1): the h planes of the tile image, one bulk copy ------------
    {
      const int seq = 2 * n, slot = seq % L::NSLOT, use = seq / L::NSLOT;
      if (use >= 1) ptx::mbar_wait(&empty[slot], (use - 1) & 1);
      tl_mark(tl, 3, n, 3);
      if (lane == 0) {
        if (gw == 0) {
          ptx::mbar_arrive_expect_tx(&full[slot], L::SLOT_BYTES);
          ptx::bulk_g2s(ring + slot * L::SLOT_BYTES, state + static_cast<int64_t>(tile) * tile_bytes(HP),
                        L::SLOT_BYTES, &full[slot]);
        } else {
          ptx::mbar_arrive(&full[slot]);
        }
      }
      __syncwarp();
      // The recurrent state was written two launches ago (the message kernel in between only reads it):
      // the first h operand is in flight before this launch waits for its predecessor's messages.
      if (n == 0) ptx::grid_dependency_wait();
    }
    // ---- x operand (k-block 0) ---------------------------------------------------------
    {
      const int seq = 2 * n + 1, slot = seq % L::NSLOT, use = seq / L::NSLOT;
      int sn[6], dn[6];
#pragma unroll
      for (int gi = 0; gi < 6; ++gi) sn[gi] = dn[gi] = 0;
      if (n + 1 < ntiles) load_idx(tile + 1, sn, dn);     // next tile's column indices, a tile ahead
      tl_mark(tl, 3, n, 0);
      float u[2][2][8], w[2][2][8];
      if (!IS_V && n > 0) k1_gather_load<true>(0, gw, lane, a.mV, si, di, u, w);   // first round trip under the slot wait
      if (use >= 1) ptx::mbar_wait(&empty[slot], (use - 1) & 1);
      tl_mark(tl, 3, n, 1);
      if (n == 0) {
        if (gw == 0) ptx::mbar_wait(boot, 0);     // tile 0: operand written by the epilogue warps (k1_boot_fill)
      } else {
        uint8_t* xs = ring + slot * L::SLOT_BYTES;
        if (IS_V) {
          k1_fill_x_v<HP>(xs, gw, lane, static_cast<int64_t>(tile) * TILE_ROWS, a.xV);
        } else {
          k1_gather_store<HP>(0, xs, gw, lane, u, w);
#pragma unroll
          for (int b = 1; b < 3; ++b) {
            k1_gather_load<true>(b, gw, lane, a.mV, si, di, u, w);
            k1_gather_store<HP>(b, xs, gw, lane, u, w);
          }
        }
        ptx::fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full[slot]);
      tl_mark(tl, 3, n, 2);
#pragma unroll
      for (int gi = 0; gi < 6; ++gi) {
        si[gi] = sn[gi];
        di[gi] = dn[gi];
      }
    }
  }
}

// ---- MMA issuer: z[128 x 256] = [x,h] . K, accumulators alternate between two TMEM halves ----
// Executed by the whole warp (converged); one elected lane issues the tcgen05 instructions.
template <int HP>
__device__ __forceinline__ void k1_mma(uint8_t* wsm, uint8_t* ring, uint64_t* bar_w,
                                       uint64_t* full, uint64_t* empty, uint64_t* acc_full, uint64_t* acc_empty,
                                       uint32_t tmem, int ntiles, long long* tl_) {
  using L = K1Smem<HP>;
  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(128, 256);
  const bool leader = ptx::elect_one();
  long long* tl = leader ? tl_ : nullptr;
  ptx::mbar_wait(bar_w, 0);          // weight images (staged by the kernel prologue)
  // descriptors of the first k-step of (ring slot 0, plane 0) and of (weight plane 0, k-block 0);
  // everything else is an offset in 16-byte units added to the low word
  const uint64_t adesc0 = ptx::umma_desc_k_nosw(ptx::smem_u32(ring), 2048, 128);
  const uint64_t bdesc0 = ptx::umma_desc_k_nosw(ptx::smem_u32(wsm), 4096, 128);
  for (int n = 0; n < ntiles; ++n) {
    const int acc = n & 1, k_use = n >> 1;
    tl_mark(tl, 2, n, 0);
    if (k_use >= 1) ptx::mbar_wait(&acc_empty[acc], (k_use - 1) & 1);
    tl_mark(tl, 2, n, 1);
    const uint32_t d_tmem = tmem + acc * 256;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // the h k-block first: its bulk copy lands long before the gathered x operand is built
      const int kb = 1 - half;
      const int seq = 2 * n + half, slot = seq % L::NSLOT, use = seq / L::NSLOT;
      ptx::mbar_wait(&full[slot], use & 1);
      tl_mark(tl, 2, n, 2 + 2 * half);
      ptx::tcgen05_fence_after();
      if (ptx::elect_one()) {
        // (A plane, B plane): small cross terms first, then hi*hi
        constexpr int NCOMB = (HP == 2) ? 3 : 1;
        const int pa_[3] = {1, 0, 0}, pb_[3] = {0, 1, 0};
        const uint64_t aslot = adesc0 + static_cast<uint32_t>((slot * L::SLOT_BYTES) >> 4);
#pragma unroll
        for (int cb = 0; cb < NCOMB; ++cb) {
          const int pa = (HP == 2) ? pa_[cb] : 0, pb = (HP == 2) ? pb_[cb] : 0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss(d_tmem, aslot + ((pa * PLANE_BYTES + k * 4096) >> 4),
                              bdesc0 + (((pb * 2 + kb) * 32768 + k * 8192) >> 4), IDESC,
                              (half | cb | k) ? 1u : 0u);
        }
        ptx::umma_commit(&empty[slot]);
        if (half == 1) ptx::umma_commit(&acc_full[acc]);
      }
      __syncwarp();
      tl_mark(tl, 2, n, 3 + 2 * half);
    }
  }
}

// ---- epilogue: thread r owns row r (TMEM lane r) of the tiles of its warpgroup -----------
// Loops are kept rolled on purpose: the epilogue is the instruction-heavy part of the step and a
// fully unrolled version (60 KB of SASS per role) was instruction-fetch bound.  The FMA-pipe math
// uses packed fp32x2 instructions (FFMA2 / FADD2 / FMUL2).

// two-pass mean / inverse standard deviation of 64 TMEM columns of this thread's row
// (nn.moments semantics); returns rstd and nmr = -mean * rstd so that LN(v) = fma(v, rstd, nmr)
__device__ __forceinline__ void row_stats64(uint32_t taddr, float& rstd, float& nmr) {
  float v[64];
  ptx::tmem_ld64(taddr, v);
  float2 s2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) s2[k] = make_float2(v[2 * k], v[2 * k + 1]);
#pragma unroll
  for (int j = 4; j < 32; ++j) s2[j & 3] = __fadd2_rn(s2[j & 3], make_float2(v[2 * j], v[2 * j + 1]));
  const float2 st = __fadd2_rn(__fadd2_rn(s2[0], s2[1]), __fadd2_rn(s2[2], s2[3]));
  const float m = (st.x + st.y) * (1.0f / 64);
  const float2 nm2 = make_float2(-m, -m);
  float2 q2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) q2[k] = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float2 t = __fadd2_rn(make_float2(v[2 * j], v[2 * j + 1]), nm2);
    q2[j & 3] = __ffma2_rn(t, t, q2[j & 3]);
  }
  const float2 qt = __fadd2_rn(__fadd2_rn(q2[0], q2[1]), __fadd2_rn(q2[2], q2[3]));
  rstd = rsqrtf((qt.x + qt.y) * (1.0f / 64) + LN_EPS);
  nmr = -m * rstd;
}

// Gate pre-activations arrive already centred: the LSTM weight images are K.C with
// C = blockdiag(I - 11^T/64) per gate (tspgnn_set_params), so the tensor core subtracts each gate's
// row mean and the LayerNorm statistics reduce to one pass of sum of squares.
__device__ __forceinline__ float row_rstd_centered64(uint32_t taddr) {
  float v[64];
  ptx::tmem_ld64(taddr, v);
  float2 q2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = make_float2(v[2 * k], v[2 * k + 1]);
    q2[k] = __fmul2_rn(t, t);
  }
#pragma unroll
  for (int j = 4; j < 32; ++j) {
    const float2 t = make_float2(v[2 * j], v[2 * j + 1]);
    q2[j & 3] = __ffma2_rn(t, t, q2[j & 3]);
  }
  const float2 qt = __fadd2_rn(__fadd2_rn(q2[0], q2[1]), __fadd2_rn(q2[2], q2[3]));
  return rsqrtf((qt.x + qt.y) * (1.0f / 64) + LN_EPS);
}

__device__ __forceinline__ float rstd_centered64_regs(const float (&v)[64]) {
  float2 q2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = make_float2(v[2 * k], v[2 * k + 1]);
    q2[k] = __fmul2_rn(t, t);
  }
#pragma unroll
  for (int j = 4; j < 32; ++j) {
    const float2 t = make_float2(v[2 * j], v[2 * j + 1]);
    q2[j & 3] = __ffma2_rn(t, t, q2[j & 3]);
  }
  const float2 qt = __fadd2_rn(__fadd2_rn(q2[0], q2[1]), __fadd2_rn(q2[2], q2[3]));
  return rsqrtf((qt.x + qt.y) * (1.0f / 64) + LN_EPS);
}

// 1 + 2^t for two values (t = -x log2 e, so this is 1 + exp(-x)).  CLAMP bounds the exponent so that
// the product of two results stays finite; tspgnn_set_params turns it off when the LayerNorm
// parameters bound |t| well below that (|LN(z)| <= sqrt(63) for 64 features).
template <bool CLAMP>
__device__ __forceinline__ float2 one_plus_ex2(float2 t) {
  if (CLAMP) t = make_float2(fminf(t.x, 60.f), fminf(t.y, 60.f));
  return __fadd2_rn(make_float2(ptx::ex2_approx(t.x), ptx::ex2_approx(t.y)), make_float2(1.0f, 1.0f));
}

// One tile of the LayerNorm-LSTM epilogue for the thread that owns row r (TMEM lane r).
// FUSED = false: the stand-alone cell kernel (K1) - the accumulator goes back to the MMA warp with the
//                last TMEM read.
// FUSED = true : the fused timestep kernel (tc_fused.cuh) - the new h planes are ALSO stored into the
//                tile's shared-memory slot `hslot_s` as the A operand of the message MLP, and the
//                accumulator is kept (the MLP layers reuse its columns).
// ln_s: shared-memory copy of this cell's LayerNorm parameters, gamma[g][j] at ln_s + (g*64+j)*4,
// beta at +1280.  (Run-time indexed constant-bank loads cost ~10 cycles each; a broadcast
// LDS.128 brings four values in ~2.)
template <int HP, int CELL, bool CLAMP, bool FUSED>
__device__ __forceinline__ void k1_cell_tile(uint8_t* gtile, int r, int lane, uint32_t t_acc, uint64_t* acc_full_bar,
                                             uint32_t acc_parity, uint64_t* acc_empty_bar, uint32_t ln_s,
                                             const float* __restrict__ vdeg_row, uint32_t hslot_s, long long* tl, int e,
                                             int n) {
  {
    float4* cg = reinterpret_cast<float4*>(gtile + HP * PLANE_BYTES) + r;   // chunk q at cg[q * 128]
    uint4* hg = reinterpret_cast<uint4*>(gtile) + r;                        // plane p, chunk ch at hg[p*1024 + ch*128]
    // first 32 columns of the old cell state: issued before the accumulator is ready
    float4 cur[4], nx1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cur[q] = cg[q * 128];
      nx1[q] = cg[(4 + q) * 128];
    }
    tl_mark(tl, e, n, 0);
    ptx::mbar_wait(acc_full_bar, acc_parity);
    tl_mark(tl, e, n, 1);
    ptx::tcgen05_fence_after();

    if (CELL == 0 && vdeg_row != nullptr) {
      // folded E_msg_V output layer: z += deg(v) * (b4 . Kx)   (vertex tiles only, 5 % of the rows).
      // Fully unrolled so that the bias values are constant-bank operands of the FMAs.
      const float dg = *vdeg_row;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float zz[64];
        ptx::tmem_ld64(t_acc + g * 64, zz);
#pragma unroll
        for (int i = 0; i < 64; ++i) zz[i] = fmaf(dg, c_vfold_bias[g * 64 + i], zz[i]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float w16[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w16[i] = zz[q * 16 + i];
          ptx::tmem_st16(t_acc + g * 64 + q * 16, w16);
        }
      }
    }

    // ---- LayerNorm statistics of the four gates ------------------------------------------
    float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      const float rstd = row_rstd_centered64(t_acc + g * 64);
      if (g == 0) rs0 = rstd;
      else if (g == 1) rs1 = rstd;
      else if (g == 2) rs2 = rstd;
      else rs3 = rstd;
    }
    tl_mark(tl, e, n, 2);
    // ---- new cell state before its LayerNorm; parked in the (consumed) i-gate columns -------
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      float4 nx2[4];
      const int cn_ = (cc < 2) ? cc + 2 : cc;            // prefetch the 16 columns of c needed two iterations ahead
#pragma unroll
      for (int q = 0; q < 4; ++q) nx2[q] = cg[(cn_ * 4 + q) * 128];
      float vi[16], vj[16], vf[16], cnew[16];
      float4 gq[6];
      ptx::tmem_ld16x3(t_acc + 0 * 64 + cc * 16, t_acc + 1 * 64 + cc * 16, t_acc + 2 * 64 + cc * 16, vi, vj, vf);
      const uint32_t lcc = ln_s + cc * 64;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        float4 Gi, Bi, Gj, Bj, Gf, Bf;
        if ((p & 1) == 0) {
          Gi = ptx::lds128f(lcc + 0 * 256 + (p >> 1) * 16);
          Bi = ptx::lds128f(lcc + 1280 + 0 * 256 + (p >> 1) * 16);
          Gj = ptx::lds128f(lcc + 1 * 256 + (p >> 1) * 16);
          Bj = ptx::lds128f(lcc + 1280 + 1 * 256 + (p >> 1) * 16);
          Gf = ptx::lds128f(lcc + 2 * 256 + (p >> 1) * 16);
          Bf = ptx::lds128f(lcc + 1280 + 2 * 256 + (p >> 1) * 16);
          gq[0] = Gi; gq[1] = Bi; gq[2] = Gj; gq[3] = Bj; gq[4] = Gf; gq[5] = Bf;
        }
        const bool hiq = (p & 1) != 0;
        const float2 gi = hiq ? make_float2(gq[0].z, gq[0].w) : make_float2(gq[0].x, gq[0].y);
        const float2 bi = hiq ? make_float2(gq[1].z, gq[1].w) : make_float2(gq[1].x, gq[1].y);
        const float2 gj = hiq ? make_float2(gq[2].z, gq[2].w) : make_float2(gq[2].x, gq[2].y);
        const float2 bj = hiq ? make_float2(gq[3].z, gq[3].w) : make_float2(gq[3].x, gq[3].y);
        const float2 gf = hiq ? make_float2(gq[4].z, gq[4].w) : make_float2(gq[4].x, gq[4].y);
        const float2 bf = hiq ? make_float2(gq[5].z, gq[5].w) : make_float2(gq[5].x, gq[5].y);
        const float2 in = __ffma2_rn(__fmul2_rn(make_float2(vi[2 * p], vi[2 * p + 1]), make_float2(rs0, rs0)), gi, bi);
        const float2 jn = __ffma2_rn(__fmul2_rn(make_float2(vj[2 * p], vj[2 * p + 1]), make_float2(rs1, rs1)), gj, bj);
        const float2 fn = __ffma2_rn(__fmul2_rn(make_float2(vf[2 * p], vf[2 * p + 1]), make_float2(rs2, rs2)), gf, bf);
        const float4 c4 = cur[p >> 1];
        const float2 cold = (p & 1) ? make_float2(c4.z, c4.w) : make_float2(c4.x, c4.y);
        // c*sigmoid(f) + sigmoid(i)*relu(j) = (c*Q + relu(j)*P) / (P*Q), P = 1+e^-f, Q = 1+e^-i:
        // one reciprocal for the two logistic functions
        const float2 P = one_plus_ex2<CLAMP>(fn), Q = one_plus_ex2<CLAMP>(in);   // in / fn are already -x log2 e
        const float2 den = __fmul2_rn(P, Q);
        const float2 num = __ffma2_rn(cold, Q, __fmul2_rn(ptx::relu2(jn), P));
        const float2 cn2 = __fmul2_rn(num, make_float2(ptx::rcp_approx(den.x), ptx::rcp_approx(den.y)));
        cnew[2 * p] = cn2.x;
        cnew[2 * p + 1] = cn2.y;
      }
      ptx::tmem_st16(t_acc + cc * 16, cnew);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        cur[q] = nx1[q];
        nx1[q] = nx2[q];
      }
    }
    tl_mark(tl, e, n, 3);
    float crs, cm;
    row_stats64(t_acc, crs, cm);
    tl_mark(tl, e, n, 4);
    // ---- LayerNorm of the cell state, output gate, new h; straight to global memory -------------
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      float vo[16], cs[16];
      ptx::tmem_ld16x2(t_acc + 3 * 64 + cc * 16, t_acc + cc * 16, vo, cs);
      if (!FUSED && cc == 3) {     // last TMEM read of this tile: hand the accumulator back to the MMA warp
        ptx::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_empty_bar);
      }
      const uint32_t lcc = ln_s + cc * 64;
      float2 c2[8];
      uint32_t hi[8], lo[8];
      float4 gq[4];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        if ((p & 1) == 0) {
          gq[0] = ptx::lds128f(lcc + 3 * 256 + (p >> 1) * 16);
          gq[1] = ptx::lds128f(lcc + 1280 + 3 * 256 + (p >> 1) * 16);
          gq[2] = ptx::lds128f(lcc + 4 * 256 + (p >> 1) * 16);
          gq[3] = ptx::lds128f(lcc + 1280 + 4 * 256 + (p >> 1) * 16);
        }
        const bool hiq = (p & 1) != 0;
        const float2 go = hiq ? make_float2(gq[0].z, gq[0].w) : make_float2(gq[0].x, gq[0].y);
        const float2 bo = hiq ? make_float2(gq[1].z, gq[1].w) : make_float2(gq[1].x, gq[1].y);
        const float2 gs = hiq ? make_float2(gq[2].z, gq[2].w) : make_float2(gq[2].x, gq[2].y);
        const float2 bs = hiq ? make_float2(gq[3].z, gq[3].w) : make_float2(gq[3].x, gq[3].y);
        const float2 on = __ffma2_rn(__fmul2_rn(make_float2(vo[2 * p], vo[2 * p + 1]), make_float2(rs3, rs3)), go, bo);
        c2[p] = __ffma2_rn(__ffma2_rn(make_float2(cs[2 * p], cs[2 * p + 1]), make_float2(crs, crs),
                                      make_float2(cm, cm)), gs, bs);
        const float2 eo = one_plus_ex2<false>(on);   // on is already -x log2 e; 2^t = inf gives h = 0, the right limit
        const float2 hn = __fmul2_rn(ptx::relu2(c2[p]), make_float2(ptx::rcp_approx(eo.x), ptx::rcp_approx(eo.y)));
        ptx::split_bf16x2_p(hn, hi[p], lo[p]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        cg[(cc * 4 + q) * 128] = make_float4(c2[2 * q].x, c2[2 * q].y, c2[2 * q + 1].x, c2[2 * q + 1].y);
#pragma unroll
      for (int q = 0; q < 2; ++q) {   // two 16-B chunks of 8 bf16
        const uint4 h4 = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
        const uint4 l4 = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        hg[(cc * 2 + q) * 128] = h4;
        if (HP == 2) hg[1024 + (cc * 2 + q) * 128] = l4;
        if (FUSED) {     // same chunk-major image in shared memory: A operand of the first MLP layer
          const uint32_t so = hslot_s + (cc * 2 + q) * 2048 + r * 16;
          ptx::sts128(so, h4);
          if (HP == 2) ptx::sts128(so + PLANE_BYTES, l4);
        }
      }
    }
    tl_mark(tl, e, n, 5);
  }
}

// Stand-alone cell kernel (K1), second form of the tile epilogue.  Differences from k1_cell_tile:
//  * the statistics of the cell-state LayerNorm are accumulated in the gate pass (shifted one-pass sums
//    around the row's first value instead of a separate two-pass read of the parked c~ row);
//  * after the gate pass the o-gate row and the parked c~ row are copied into registers (128 values) and
//    the accumulator goes back to the MMA warp at once: the MMAs of this warpgroup's next tile (3.2 k
//    cycles of tensor pipe) run under the final pass instead of after it.  With two 256-column
//    accumulators in TMEM that wait was 3 of every 12.5 k cycles of a warpgroup;
//  * the gate LayerNorm statistics load the next gate's 64 columns while the current ones are squared.
// The final pass is unrolled (its operands are register arrays).
// (Measured and dropped: ONE reciprocal for all three logistic gates -- R = 1/(P.Q.E), c~ = num.R.E,
// sigmoid(o) = R.P.Q, four transcendentals per column instead of five.  It shortens the epilogue by 1.5 k
// cycles per tile but moves the o-gate work in front of the accumulator hand-back: the final pass (1.5 k)
// no longer covers the next tile's MMAs (3.2 k) and the tile rate drops, K1 35.0 vs 31.2 us.  The pace of
// this kernel is max((MMA + gate passes) / 2 accumulators, all passes / 2 warpgroups); both are ~5.2 k.)

template <int HP, int CELL, bool CLAMP>
__device__ __forceinline__ void k1_cell_tile2(uint8_t* gtile, int r, int lane, uint32_t t_acc, uint64_t* acc_full_bar,
                                              uint32_t acc_parity, uint64_t* acc_empty_bar, uint32_t ln_s,
                                              const float* __restrict__ vdeg_row, long long* tl, int e, int n) {
  float4* cg = reinterpret_cast<float4*>(gtile + HP * PLANE_BYTES) + r;   // chunk q at cg[q * 128]
  uint4* hg = reinterpret_cast<uint4*>(gtile) + r;                        // plane p, chunk ch at hg[p*1024 + ch*128]
  // first 16 columns of the old cell state: issued before the accumulator is ready
  float4 cur[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) cur[q] = cg[q * 128];
  tl_mark(tl, e, n, 0);
  ptx::mbar_wait(acc_full_bar, acc_parity);
  tl_mark(tl, e, n, 1);
  ptx::tcgen05_fence_after();

  if (CELL == 0 && vdeg_row != nullptr) {
    // folded E_msg_V output layer: z += deg(v) * (b4 . Kx)   (vertex tiles only, 5 % of the rows)
    const float dg = *vdeg_row;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float zz[64];
      ptx::tmem_ld64(t_acc + g * 64, zz);
#pragma unroll
      for (int i = 0; i < 64; ++i) zz[i] = fmaf(dg, c_vfold_bias[g * 64 + i], zz[i]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float w16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w16[i] = zz[q * 16 + i];
        ptx::tmem_st16(t_acc + g * 64 + q * 16, w16);
      }
    }
  }

  // ---- LayerNorm statistics of the four gates (centred by the weights: sum of squares only) ----
  float rs0, rs1, rs2, rs3;
  {
    float va[64], vb[64];
    ptx::tmem_ld64_nowait(t_acc, va);
    ptx::tmem_ld64_nowait(t_acc + 64, vb);
    ptx::tmem_wait_ld();
    rs0 = rstd_centered64_regs(va);
    ptx::tmem_ld64_nowait(t_acc + 128, va);
    rs1 = rstd_centered64_regs(vb);
    ptx::tmem_wait_ld();
    ptx::tmem_ld64_nowait(t_acc + 192, vb);
    rs2 = rstd_centered64_regs(va);
    ptx::tmem_wait_ld();
    rs3 = rstd_centered64_regs(vb);
  }
  tl_mark(tl, e, n, 2);
  // ---- new cell state before its LayerNorm; parked in the (consumed) i-gate columns -------
  float kshift = 0.f;
  float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int cc = 0; cc < 4; ++cc) {
    float4 nx[4];
    const int cn_ = (cc < 3) ? cc + 1 : cc;            // the 16 columns of c of the next iteration
#pragma unroll
    for (int q = 0; q < 4; ++q) nx[q] = cg[(cn_ * 4 + q) * 128];
    float vi[16], vj[16], vf[16];
    float4 gq[6];
    ptx::tmem_ld16x3(t_acc + 0 * 64 + cc * 16, t_acc + 1 * 64 + cc * 16, t_acc + 2 * 64 + cc * 16, vi, vj, vf);
    const uint32_t lcc = ln_s + cc * 64;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if ((p & 1) == 0) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          gq[2 * g] = ptx::lds128f(lcc + g * 256 + (p >> 1) * 16);
          gq[2 * g + 1] = ptx::lds128f(lcc + 1280 + g * 256 + (p >> 1) * 16);
        }
      }
      const bool hiq = (p & 1) != 0;
      const float2 gi = hiq ? make_float2(gq[0].z, gq[0].w) : make_float2(gq[0].x, gq[0].y);
      const float2 bi = hiq ? make_float2(gq[1].z, gq[1].w) : make_float2(gq[1].x, gq[1].y);
      const float2 gj = hiq ? make_float2(gq[2].z, gq[2].w) : make_float2(gq[2].x, gq[2].y);
      const float2 bj = hiq ? make_float2(gq[3].z, gq[3].w) : make_float2(gq[3].x, gq[3].y);
      const float2 gf = hiq ? make_float2(gq[4].z, gq[4].w) : make_float2(gq[4].x, gq[4].y);
      const float2 bf = hiq ? make_float2(gq[5].z, gq[5].w) : make_float2(gq[5].x, gq[5].y);
      const float2 in = __ffma2_rn(__fmul2_rn(make_float2(vi[2 * p], vi[2 * p + 1]), make_float2(rs0, rs0)), gi, bi);
      const float2 jn = __ffma2_rn(__fmul2_rn(make_float2(vj[2 * p], vj[2 * p + 1]), make_float2(rs1, rs1)), gj, bj);
      const float2 fn = __ffma2_rn(__fmul2_rn(make_float2(vf[2 * p], vf[2 * p + 1]), make_float2(rs2, rs2)), gf, bf);
      const float4 c4 = cur[p >> 1];
      const float2 cold = (p & 1) ? make_float2(c4.z, c4.w) : make_float2(c4.x, c4.y);
      // c*sigmoid(f) + sigmoid(i)*relu(j) = (c*Q + relu(j)*P) / (P*Q), P = 1+e^-f, Q = 1+e^-i:
      // one reciprocal for the two logistic functions
      const float2 P = one_plus_ex2<CLAMP>(fn), Q = one_plus_ex2<CLAMP>(in);   // in / fn are already -x log2 e
      const float2 den = __fmul2_rn(P, Q);
      const float2 num = __ffma2_rn(cold, Q, __fmul2_rn(ptx::relu2(jn), P));
      const float2 cn2 = __fmul2_rn(num, make_float2(ptx::rcp_approx(den.x), ptx::rcp_approx(den.y)));
      vi[2 * p] = cn2.x;
      vi[2 * p + 1] = cn2.y;
      if (cc == 0 && p == 0) kshift = cn2.x;
      const float2 t = __fadd2_rn(cn2, make_float2(-kshift, -kshift));
      s1 = __fadd2_rn(s1, t);
      s2 = __ffma2_rn(t, t, s2);
    }
    ptx::tmem_st16_nowait(t_acc + cc * 16, vi);
#pragma unroll
    for (int q = 0; q < 4; ++q) cur[q] = nx[q];
  }
  tl_mark(tl, e, n, 3);
  // ---- o-gate row and c~ row into registers; the accumulator goes back to the MMA warp ----------
  float vo[64], cs[64];
  ptx::tmem_wait_st();
  ptx::tmem_ld64_nowait(t_acc + 192, vo);
  ptx::tmem_ld64_nowait(t_acc, cs);
  ptx::tmem_wait_ld();
  ptx::tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(acc_empty_bar);
  // nn.moments of c~ from the shifted sums: mean = k + S1/64, var = S2/64 - (S1/64)^2
  const float m1 = (s1.x + s1.y) * (1.0f / 64);
  const float var = fmaxf(fmaf(-m1, m1, (s2.x + s2.y) * (1.0f / 64)), 0.f);
  const float crs = rsqrtf(var + LN_EPS);
  const float cm = -(kshift + m1) * crs;
  tl_mark(tl, e, n, 4);
  // ---- LayerNorm of the cell state, output gate, new h; straight to global memory -------------
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const uint32_t lcc = ln_s + cc * 64;
    float2 c2[8];
    uint32_t hi[8], lo[8];
    float4 gq[4];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if ((p & 1) == 0) {
        gq[0] = ptx::lds128f(lcc + 3 * 256 + (p >> 1) * 16);
        gq[1] = ptx::lds128f(lcc + 1280 + 3 * 256 + (p >> 1) * 16);
        gq[2] = ptx::lds128f(lcc + 4 * 256 + (p >> 1) * 16);
        gq[3] = ptx::lds128f(lcc + 1280 + 4 * 256 + (p >> 1) * 16);
      }
      const bool hiq = (p & 1) != 0;
      const float2 go = hiq ? make_float2(gq[0].z, gq[0].w) : make_float2(gq[0].x, gq[0].y);
      const float2 bo = hiq ? make_float2(gq[1].z, gq[1].w) : make_float2(gq[1].x, gq[1].y);
      const float2 gs = hiq ? make_float2(gq[2].z, gq[2].w) : make_float2(gq[2].x, gq[2].y);
      const float2 bs = hiq ? make_float2(gq[3].z, gq[3].w) : make_float2(gq[3].x, gq[3].y);
      const int j = cc * 16 + 2 * p;
      const float2 on = __ffma2_rn(__fmul2_rn(make_float2(vo[j], vo[j + 1]), make_float2(rs3, rs3)), go, bo);
      c2[p] = __ffma2_rn(__ffma2_rn(make_float2(cs[j], cs[j + 1]), make_float2(crs, crs), make_float2(cm, cm)), gs, bs);
      const float2 eo = one_plus_ex2<false>(on);   // on is already -x log2 e; 2^t = inf gives h = 0, the right limit
      const float2 hn = __fmul2_rn(ptx::relu2(c2[p]), make_float2(ptx::rcp_approx(eo.x), ptx::rcp_approx(eo.y)));
      ptx::split_bf16x2_p(hn, hi[p], lo[p]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      cg[(cc * 4 + q) * 128] = make_float4(c2[2 * q].x, c2[2 * q].y, c2[2 * q + 1].x, c2[2 * q + 1].y);
#pragma unroll
    for (int q = 0; q < 2; ++q) {   // two 16-B chunks of 8 bf16
      hg[(cc * 2 + q) * 128] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
      if (HP == 2) hg[1024 + (cc * 2 + q) * 128] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
    }
  }
  tl_mark(tl, e, n, 5);
}

template <int HP, int CELL, bool CLAMP>
__device__ __forceinline__ void k1_epilogue(uint8_t* state, int t0, int ntiles, uint32_t tmem, uint64_t* acc_full,
                                            uint64_t* acc_empty, int warp, int lane, long long* tl_, uint32_t ln_s,
                                            const float* __restrict__ vdeg) {
  const int e = warp >> 2, q4 = warp & 3;
  long long* tl = (q4 == 0 && lane == 0) ? tl_ : nullptr;
  const int r = q4 * 32 + lane;
  const uint32_t t_acc = tmem + (static_cast<uint32_t>(q4 * 32) << 16) + e * 256;
  int use = 0;
  for (int n = e; n < ntiles; n += 2, ++use) {
    uint8_t* gtile = state + static_cast<int64_t>(t0 + n) * tile_bytes(HP);
    const float* vdeg_row = (vdeg != nullptr) ? vdeg + static_cast<int64_t>(t0 + n) * TILE_ROWS + r : nullptr;
    k1_cell_tile2<HP, CELL, CLAMP>(gtile, r, lane, t_acc, &acc_full[e], use & 1, &acc_empty[e], ln_s, vdeg_row, tl, e, n);
  }
}

template <int HP>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_lnlstm_kernel(const K1Args a) {
  using L = K1Smem<HP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* wsm = smem;
  uint8_t* ring = smem + L::RING_OFF;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* full = bar_w + 1;
  uint64_t* empty = full + L::NSLOT;
  uint64_t* acc_full = empty + L::NSLOT;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* boot = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(boot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  tl_gmark(a.timeline, a.tl_slot, 0);
  const bool is_v = static_cast<int>(blockIdx.x) >= a.e_ctas;
  int t0, t1;
  if (is_v) tile_range(blockIdx.x - a.e_ctas, gridDim.x - a.e_ctas, a.tilesV, t0, t1);
  else tile_range(blockIdx.x, a.e_ctas, a.tilesE, t0, t1);
  uint8_t* state = is_v ? a.stateV : a.stateE;
  const int ntiles = t1 - t0;

  if (tid == 0) {
    ptx::mbar_init(bar_w, 1);
    ptx::mbar_init(boot, 8);
    for (int s = 0; s < L::NSLOT; ++s) {
      ptx::mbar_init(&full[s], NUM_GATHER_WARPS);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], 4);
    }
    ptx::fence_mbar_init();
    if (ntiles > 0) {   // weight images: parameters, not produced by the preceding kernel
      const uint8_t* wimg = is_v ? a.wV : a.wE;
      ptx::mbar_arrive_expect_tx(bar_w, L::W_BYTES);
      for (int off = 0; off < L::W_BYTES; off += 32768) ptx::bulk_g2s(wsm + off, wimg + off, 32768, bar_w);
    }
  }
  if (warp == 8) ptx::tmem_alloc(tmem_slot, 512);
  {
    // LayerNorm parameters of this CTA's cell.  The three gates that only feed a logistic function
    // (input 0, forget 2, output 3) are stored pre-multiplied by -log2(e), with the forget bias
    // folded into beta: the epilogue then gets the exponent of 2^(-x log2 e) straight from the FMA.
    // (tspgnn_set_params builds that table once; a coalesced copy here instead of run-time indexed
    // constant-bank reads, which serialise 32-way and sit on the critical path of the last SM to start)
    float* ln_sm = reinterpret_cast<float*>(smem + L::LN_OFF);
    const float* tab = a.ln_tab + (is_v ? 0 : 2 * 5 * D);
    for (int i = tid; i < 2 * 5 * D; i += TC_THREADS) ln_sm[i] = tab[i];
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t ln_s = ptx::smem_u32(smem + L::LN_OFF);
  // everything above overlapped the tail of the previous kernel (programmatic dependent launch).  Every
  // role executes griddepcontrol.wait itself, as late as it can: only the messages (mV / xV) come from the
  // preceding kernel, so column indices, the first h operand and the first cell-state chunk are requested
  // before it.
  tl_gmark(a.timeline, a.tl_slot, 1);
  ptx::grid_launch_dependents();

  if (warp < 8) {
    ptx::setmaxnreg_inc<192>();   // ... 256 x (192 - 168) = 6144 taken by the two epilogue warpgroups
    if (ntiles > 0) {
      uint8_t* x0 = ring + (1 % L::NSLOT) * L::SLOT_BYTES;       // ring slot of sequence number 1 = x operand of tile 0
      if (is_v) k1_boot_fill<HP, true>(a, x0, warp, lane, t0);   // (waits for the preceding kernel inside)
      else k1_boot_fill<HP, false>(a, x0, warp, lane, t0);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(boot);
    } else {
      ptx::grid_dependency_wait();
    }
    tl_gmark(a.timeline, a.tl_slot, 2);
    const bool clamp = (is_v ? a.clampV : a.clampE) != 0;
    if (is_v) {
      if (clamp) k1_epilogue<HP, 0, true>(state, t0, ntiles, tmem, acc_full, acc_empty, warp, lane, a.timeline, ln_s, a.vdeg);
      else k1_epilogue<HP, 0, false>(state, t0, ntiles, tmem, acc_full, acc_empty, warp, lane, a.timeline, ln_s, a.vdeg);
    } else {
      if (clamp) k1_epilogue<HP, 1, true>(state, t0, ntiles, tmem, acc_full, acc_empty, warp, lane, a.timeline, ln_s, a.vdeg);
      else k1_epilogue<HP, 1, false>(state, t0, ntiles, tmem, acc_full, acc_empty, warp, lane, a.timeline, ln_s, a.vdeg);
    }
  } else {
    ptx::setmaxnreg_dec<120>();   // 128 x (168 - 120) = 6144 registers back to the CTA pool ...
    if (warp == 8) {
      ptx::grid_dependency_wait();
      if (ntiles > 0)
        k1_mma<HP>(wsm, ring, bar_w, full, empty, acc_full, acc_empty, tmem, ntiles, a.timeline);
    } else {
      if (ntiles == 0) ptx::grid_dependency_wait();
      if (is_v) k1_producer<HP, true>(a, state, ring, full, empty, boot, t0, ntiles, warp - 9, lane);
      else k1_producer<HP, false>(a, state, ring, full, empty, boot, t0, ntiles, warp - 9, lane);
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  tl_gmark(a.timeline, a.tl_slot, 3);
  if (warp == 8) ptx::tmem_dealloc(tmem, 512);
}

// ====================================================================================
// K2: message MLP chain (+ scatter / message store / vote)
// ====================================================================================
// 448 threads: three chain warpgroups (warps 0-3, 4-7, 8-11; tile n -> chain n % 3), warp 12 issues
// the MMAs and owns the TMEM allocation (3 x 64 columns), warp 13 prefetches tile inputs.
// A tile lives in ONE 32 KB shared-memory slot for its whole chain: the h planes arrive by bulk
// copy, every hidden layer's activations overwrite them in place once that layer's MMA has
// completed (the accumulator-full barrier is exactly that event), and the fp32 messages are staged
// there for the scatter.  Five slots: three tiles in flight plus two prefetched.
constexpr int K2_CHAINS = 3;          // four chains (576 threads, 96 registers) measured the same: K2 22.8 vs 22.6 us
constexpr int K2_THREADS = 128 * K2_CHAINS + 64;
constexpr int K2_MMA_WARP = 4 * K2_CHAINS;

template <int HP>
struct K2Smem {
  static constexpr int W_BYTES = 4 * HP * 8192;            // [layer][plane] 8 KB images
  static constexpr int IN_BYTES = HP * PLANE_BYTES;        // h planes of one tile
  static constexpr int SLOT_BYTES = 32768;                 // >= IN_BYTES, = fp32 staging of 128 x 64 messages
  static constexpr int NSLOT = 5;
  static constexpr int SLOT_OFF = W_BYTES;
  static constexpr int BAR_OFF = SLOT_OFF + NSLOT * SLOT_BYTES;
  static constexpr int NBAR = 1 + 2 * NSLOT + 2 * K2_CHAINS;
  static constexpr int BIAS_OFF = (BAR_OFF + 8 * NBAR + 16 + 15) & ~15;   // b[4][64] of this CTA's MLP
  static constexpr int TOTAL = BIAS_OFF + 4 * D * 4;
  static constexpr int DYN_BYTES = TOTAL + 128;
};
static_assert(K2Smem<2>::DYN_BYTES <= 232448, "K2 shared memory budget (227 KB)");

// fp32 staging of a 128 x 64 message tile: 256-B rows, 16-B chunk index XOR (row & 7)
__device__ __forceinline__ uint32_t stage_off(int r, int chunk) {
  return static_cast<uint32_t>(r * 256 + ((chunk ^ (r & 7)) << 4));
}

// ---- one warpgroup: epilogues of the layer chain of its tiles ------------------------------
// ROLE 0 = V rows (V_msg_E, message stored), 1 = E rows (E_msg_V, scatter-add), 2 = E rows vote,
// 3 = E rows with the output layer folded into the vertex cell (three layers, scatter-add of a3)
template <int HP, int ROLE>
__device__ __forceinline__ void k2_chain(const K2Args& a, uint8_t* slots, uint64_t* acc_full, uint64_t* act_ready,
                                         uint64_t* slot_free, uint32_t tmem, int t0, int ntiles, int warp, int lane,
                                         uint32_t bias_s) {
  using L = K2Smem<HP>;
  constexpr int NL = (ROLE >= 2) ? 3 : 4;
  const int e = warp >> 2, q4 = warp & 3;
  const int r = q4 * 32 + lane;
  const uint32_t t_acc = tmem + (static_cast<uint32_t>(q4 * 32) << 16) + e * 64;
  const int64_t n_rows = (ROLE == 0) ? a.nV : a.nE;
  long long* tl = (q4 == 0 && lane == 0 && e < 2) ? a.timeline : nullptr;
  const uint32_t slots_s = ptx::smem_u32(slots);
  uint32_t step = 0;
  for (int n = e; n < ntiles; n += K2_CHAINS) {
    const int64_t row0 = static_cast<int64_t>(t0 + n) * TILE_ROWS;
    const int slot = n % L::NSLOT;
    const uint32_t b_s = slots_s + slot * L::SLOT_BYTES;
    // this half-warp's 32 pairs of the tile's scatter plan (lane c16 holds pairs c16 and c16 + 16), fetched a
    // whole layer chain ahead of the scatter that uses them
    int e_row0 = 0, e_row1 = 0, e_v0 = -1, e_v1 = -1;
    if (ROLE == 1 || ROLE == 3) {
      const int64_t eo = static_cast<int64_t>(t0 + n) * (2 * TILE_ROWS) + (q4 * 2 + (lane >> 4)) * 32 + (lane & 15);
      e_row0 = __ldg(a.ent_row + eo);
      e_row1 = __ldg(a.ent_row + eo + 16);
      e_v0 = __ldg(a.ent_v + eo);
      e_v1 = __ldg(a.ent_v + eo + 16);
    }
    float v[64];
#pragma unroll 1
    for (int l = 0; l < NL; ++l, ++step) {
      ptx::mbar_wait(&acc_full[e], step & 1);
      tl_mark(tl, e, n, l);
      ptx::tcgen05_fence_after();
      ptx::tmem_ld64(t_acc, v);
      const bool hidden = (ROLE >= 2) || (l < 3);
      const bool feeds_mma = l < NL - 1;
      if (ROLE == 3 && !feeds_mma) {
        // last hidden layer of the folded chain: bias + ReLU only, kept in fp32 for the scatter
        const uint32_t bl = bias_s + l * 256;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 bb = ptx::lds128f(bl + q * 16);
          const float2 x0 = ptx::relu2(__fadd2_rn(make_float2(v[4 * q], v[4 * q + 1]), make_float2(bb.x, bb.y)));
          const float2 x1 = ptx::relu2(__fadd2_rn(make_float2(v[4 * q + 2], v[4 * q + 3]), make_float2(bb.z, bb.w)));
          v[4 * q] = x0.x; v[4 * q + 1] = x0.y; v[4 * q + 2] = x1.x; v[4 * q + 3] = x1.y;
        }
      } else if (hidden) {
        const uint32_t nxt = b_s + r * 16;          // in place: this layer's MMA has finished reading the slot
        const uint32_t bl = bias_s + l * 256;       // broadcast LDS.128: four bias values per load
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t hi[4], lo[4];
          const float4 ba = ptx::lds128f(bl + ch * 32), bb = ptx::lds128f(bl + ch * 32 + 16);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int j = ch * 8 + 2 * p;
            const float2 bp = (p == 0) ? make_float2(ba.x, ba.y) : (p == 1) ? make_float2(ba.z, ba.w)
                            : (p == 2) ? make_float2(bb.x, bb.y) : make_float2(bb.z, bb.w);
            const float2 x = ptx::relu2(__fadd2_rn(make_float2(v[j], v[j + 1]), bp));
            v[j] = x.x;
            v[j + 1] = x.y;
            ptx::split_bf16x2_p(x, hi[p], lo[p]);
          }
          if (feeds_mma) {
            ptx::sts128(nxt + ch * 2048, make_uint4(hi[0], hi[1], hi[2], hi[3]));
            if (HP == 2) ptx::sts128(nxt + PLANE_BYTES + ch * 2048, make_uint4(lo[0], lo[1], lo[2], lo[3]));
          }
          if (ROLE == 1 && HP == 2 && a.act_out != nullptr) {
            // coalesced: a warp writes 512 contiguous bytes per store (thread = row, 16 bytes per chunk)
            uint4* g = reinterpret_cast<uint4*>(a.act_out + (static_cast<int64_t>(t0 + n) * 3 + l) * (2 * PLANE_BYTES)) + r;
            // streaming stores: 75 MB per timestep that nobody reads before the reverse pass must not push the
            // recurrent state (54 MB) out of the L2
            __stcs(g + ch * 128, make_uint4(hi[0], hi[1], hi[2], hi[3]));
            __stcs(g + 1024 + ch * 128, make_uint4(lo[0], lo[1], lo[2], lo[3]));
          }
        }
        if (feeds_mma) ptx::fence_proxy_async_smem();
      }
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&act_ready[e]);   // activations written / accumulator drained
    }
    tl_mark(tl, e, n, 4);
    if (ROLE == 2) {
      // 64 -> 1 tail of E_vote on the fp32 layer-3 activations (model.py:107-128)
      const int64_t grow = row0 + r;
      float s = c_vote_tail.b4;
#pragma unroll
      for (int j = 0; j < 64; ++j) s = fmaf(v[j], c_vote_tail.w4[j], s);
      if (grow < n_rows) a.vote[grow] = s;
    } else {
      // stage the fp32 messages in the tile's slot (the last MMA has finished reading it); each warp
      // then walks its own 32 rows, two rows per instruction (a half-warp covers the 256 bytes of a
      // row with 16-byte accesses)
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (ROLE == 3) {
          ptx::sts128f(b_s + stage_off(r, q), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
        } else {
          const float4 bb = ptx::lds128f(bias_s + 3 * 256 + q * 16);
          ptx::sts128f(b_s + stage_off(r, q),
                       make_float4(v[4 * q] + bb.x, v[4 * q + 1] + bb.y, v[4 * q + 2] + bb.z, v[4 * q + 3] + bb.w));
        }
      }
      // vertex rows: a warp reads back its own 32 rows; edge rows: the sorted pairs of a half-warp reference
      // rows staged by any warp of the chain
      if (ROLE == 0) __syncwarp();
      else asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");
      tl_mark(tl, e, n, 5);
      const int64_t g0 = row0 + q4 * 32;
      const int hw = lane >> 4, c16 = lane & 15;
      if (ROLE == 0) {
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int rr = 2 * i + hw;
          if (g0 + rr < n_rows) {
            const float4 m = ptx::lds128f(b_s + stage_off(q4 * 32 + rr, c16));
            *reinterpret_cast<float4*>(a.mV + (g0 + rr) * D + 4 * c16) = m;
          }
        }
      } else {
        // one reduction per GROUP of (row, endpoint) pairs with the same vertex (tc_scatter_plan_kernel):
        // half-warp hwid walks pairs [32 hwid, 32 hwid + 32) of the tile's sorted list, two batches of 16
        // row chunks in flight (shared-memory latency is several hundred cycles while the MMA operand
        // fetch owns the read port), accumulates while the vertex stays the same and reduces when it changes.
        int cur_v = -1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float4 m[16];
          int vv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int rr = __shfl_sync(0xffffffffu, b ? e_row1 : e_row0, (lane & 16) | i);
            vv[i] = __shfl_sync(0xffffffffu, b ? e_v1 : e_v0, (lane & 16) | i);
            m[i] = ptx::lds128f(b_s + stage_off(rr, c16));
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool nw = vv[i] != cur_v;      // predicated, branch-free: the two half-warps change vertex at different pairs
            ptx::red_add_v4_if(nw && cur_v >= 0 && !TSPGNN_DBG_NO_RED, a.xV + static_cast<int64_t>(cur_v < 0 ? 0 : cur_v) * D + 4 * c16, acc);
            cur_v = vv[i];
            acc.x = nw ? m[i].x : acc.x + m[i].x;
            acc.y = nw ? m[i].y : acc.y + m[i].y;
            acc.z = nw ? m[i].z : acc.z + m[i].z;
            acc.w = nw ? m[i].w : acc.w + m[i].w;
          }
          if (b == 0) tl_mark(tl, e, n, 6);
        }
        if (cur_v >= 0) ptx::red_add_v4(a.xV + static_cast<int64_t>(cur_v) * D + 4 * c16, acc);
      }
    }
    // every warp of the chain is done with the slot: hand it back to the loader
    asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");
    if (q4 == 0 && lane == 0) ptx::mbar_arrive(&slot_free[slot]);
    tl_mark(tl, e, n, 7);
  }
}

// Executed by the whole warp (converged); one elected lane issues the tcgen05 instructions.
template <int HP>
__device__ __forceinline__ void k2_mma(uint8_t* wsm, uint8_t* slots, uint64_t* bar_w, uint64_t* slot_full,
                                       uint64_t* acc_full, uint64_t* act_ready, uint32_t tmem, int ntiles, int n_layers,
                                       long long* tl_) {
  using L = K2Smem<HP>;
  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(128, 64);
  const bool leader = ptx::elect_one();
  long long* tl = leader ? tl_ : nullptr;
  ptx::mbar_wait(bar_w, 0);          // weight images (staged by the kernel prologue)
  const uint64_t sdesc0 = ptx::umma_desc_k_nosw(ptx::smem_u32(slots), 2048, 128);
  const uint64_t bdesc0 = ptx::umma_desc_k_nosw(ptx::smem_u32(wsm), 1024, 128);
  const int nrounds = (ntiles + K2_CHAINS - 1) / K2_CHAINS;
  for (int rd = 0; rd < nrounds; ++rd) {
    for (int l = 0; l < n_layers; ++l) {
#pragma unroll
      for (int e = 0; e < K2_CHAINS; ++e) {
        const int n = K2_CHAINS * rd + e;
        if (n >= ntiles) continue;
        const int slot = n % L::NSLOT;
        const uint32_t step = static_cast<uint32_t>(rd * n_layers + l);
        if (l == 0) ptx::mbar_wait(&slot_full[slot], (n / L::NSLOT) & 1);
        if (step > 0) ptx::mbar_wait(&act_ready[e], (step - 1) & 1);
        ptx::tcgen05_fence_after();
        if (ptx::elect_one()) {
          const uint64_t adesc = sdesc0 + static_cast<uint32_t>((slot * L::SLOT_BYTES) >> 4);
          const uint64_t bdesc = bdesc0 + static_cast<uint32_t>((l * HP * 8192) >> 4);
          constexpr int NCOMB = (HP == 2) ? 3 : 1;
          const int pa_[3] = {1, 0, 0}, pb_[3] = {0, 1, 0};
#pragma unroll
          for (int cb = 0; cb < NCOMB; ++cb) {
            const int pa = (HP == 2) ? pa_[cb] : 0, pb = (HP == 2) ? pb_[cb] : 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_ss(tmem + e * 64, adesc + ((pa * PLANE_BYTES + k * 4096) >> 4),
                                bdesc + ((pb * 8192 + k * 2048) >> 4), IDESC, (cb | k) ? 1u : 0u);
          }
          ptx::umma_commit(&acc_full[e]);
        }
        __syncwarp();
        if (e < 2) tl_mark(tl, 2, n, l);
      }
    }
  }
}

template <int HP>
__global__ void __launch_bounds__(K2_THREADS, 1) tc_mlp_kernel(const K2Args a) {
  using L = K2Smem<HP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* wsm = smem;
  uint8_t* slots = smem + L::SLOT_OFF;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* slot_full = bar_w + 1;
  uint64_t* slot_free = slot_full + L::NSLOT;
  uint64_t* acc_full = slot_free + L::NSLOT;
  uint64_t* act_ready = acc_full + K2_CHAINS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + K2_CHAINS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  tl_gmark(a.timeline, a.tl_slot, 0);
  const bool is_v = static_cast<int>(blockIdx.x) >= a.e_ctas;
  int t0, t1;
  if (is_v) tile_range(blockIdx.x - a.e_ctas, gridDim.x - a.e_ctas, a.tilesV, t0, t1);
  else tile_range(blockIdx.x, a.e_ctas, a.tilesE, t0, t1);
  const uint8_t* state = is_v ? a.stateV : a.stateE;
  const int ntiles = t1 - t0;

  if (tid == 0) {
    ptx::mbar_init(bar_w, 1);
    for (int s = 0; s < L::NSLOT; ++s) {
      ptx::mbar_init(&slot_full[s], 1);
      ptx::mbar_init(&slot_free[s], 1);
    }
    for (int s = 0; s < K2_CHAINS; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&act_ready[s], 4);
    }
    ptx::fence_mbar_init();
    if (ntiles > 0) {   // weight images: parameters, not produced by the preceding kernel
      const uint8_t* wimg = is_v ? a.wV : a.wE;
      ptx::mbar_arrive_expect_tx(bar_w, L::W_BYTES);
      for (int off = 0; off < L::W_BYTES; off += 16384) ptx::bulk_g2s(wsm + off, wimg + off, 16384, bar_w);
    }
  }
  if (warp == K2_MMA_WARP) ptx::tmem_alloc(tmem_slot, 256);
  {
    float* bias_sm = reinterpret_cast<float*>(smem + L::BIAS_OFF);
    const float* tab = a.bias_tab + (a.vote_mode ? 2 : (is_v ? 0 : 1)) * 4 * D;
    for (int i = tid; i < 4 * D; i += K2_THREADS) bias_sm[i] = tab[i];
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t bias_s = ptx::smem_u32(smem + L::BIAS_OFF);
  tl_gmark(a.timeline, a.tl_slot, 1);
  ptx::grid_dependency_wait();       // prologue above overlaps the previous kernel's tail
  ptx::grid_launch_dependents();
  tl_gmark(a.timeline, a.tl_slot, 2);
  if (a.zero_word != nullptr && blockIdx.x == 0 && tid == 0) *a.zero_word = 0u;

  if (warp < 4 * K2_CHAINS) {
    if (a.vote_mode) k2_chain<HP, 2>(a, slots, acc_full, act_ready, slot_free, tmem, t0, ntiles, warp, lane, bias_s);
    else if (is_v) k2_chain<HP, 0>(a, slots, acc_full, act_ready, slot_free, tmem, t0, ntiles, warp, lane, bias_s);
    else if (a.fold) k2_chain<HP, 3>(a, slots, acc_full, act_ready, slot_free, tmem, t0, ntiles, warp, lane, bias_s);
    else k2_chain<HP, 1>(a, slots, acc_full, act_ready, slot_free, tmem, t0, ntiles, warp, lane, bias_s);
  } else if (warp == K2_MMA_WARP) {
    if (ntiles > 0)
      k2_mma<HP>(wsm, slots, bar_w, slot_full, acc_full, act_ready, tmem, ntiles,
                 (a.vote_mode || (a.fold && !is_v)) ? 3 : 4, a.timeline);
  } else if (warp == K2_MMA_WARP + 1) {
    if (lane == 0) {
      for (int n = 0; n < ntiles; ++n) {
        const int slot = n % L::NSLOT, use = n / L::NSLOT;
        if (use >= 1) ptx::mbar_wait(&slot_free[slot], (use - 1) & 1);
        ptx::mbar_arrive_expect_tx(&slot_full[slot], L::IN_BYTES);
        ptx::bulk_g2s(slots + slot * L::SLOT_BYTES, state + static_cast<int64_t>(t0 + n) * tile_bytes(HP), L::IN_BYTES,
                      &slot_full[slot]);
      }
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  tl_gmark(a.timeline, a.tl_slot, 3);
  if (warp == K2_MMA_WARP) ptx::tmem_dealloc(tmem, 256);
}

// Scatter plan of the message kernel: EV^T . msg adds every edge row to its two endpoint vertices
// (instance_loader.py:63-66).  An SM retires about one 256-byte row reduction per 12.5 cycles
// (tools/microbench/bulk_red.cu), so one reduction per (row, endpoint) costs 3.3 k cycles per tile.  Sorted
// by vertex, the 256 (row, vertex) pairs of a tile form ~33 groups for a complete graph (every vertex of an
// instance is touched ~8 times per tile): the message kernel sums a group in registers and issues ONE
// reduction for it.  One CTA per tile, thread i = pair i (i < 128: src of row i, else dst of row i - 128);
// rank by counting (256 keys).  Generic: any graph, any edge order.
__global__ void __launch_bounds__(2 * TILE_ROWS) tc_scatter_plan_kernel(const int32_t* __restrict__ src,
                                                                        const int32_t* __restrict__ dst, int64_t n_edges,
                                                                        uint8_t* __restrict__ ent_row,
                                                                        int32_t* __restrict__ ent_v) {
  __shared__ int key[2 * TILE_ROWS];
  const int i = threadIdx.x, row = i & (TILE_ROWS - 1);
  const int64_t e = static_cast<int64_t>(blockIdx.x) * TILE_ROWS + row;
  int v = 0x7fffffff;
  if (e < n_edges) v = (i < TILE_ROWS) ? src[e] : dst[e];
  key[i] = v;
  __syncthreads();
  int rank = 0;
#pragma unroll 8
  for (int j = 0; j < 2 * TILE_ROWS; ++j) {
    const int kj = key[j];
    rank += (kj < v || (kj == v && j < i)) ? 1 : 0;
  }
  const int64_t o = static_cast<int64_t>(blockIdx.x) * (2 * TILE_ROWS) + rank;
  ent_row[o] = static_cast<uint8_t>(row);
  ent_v[o] = (v == 0x7fffffff) ? -1 : v;
}

// deg[v] = number of edge rows incident to vertex v (K1Args::vdeg); `deg` must be zeroed first
__global__ void __launch_bounds__(256) tc_degree_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                        int64_t n_edges, float* __restrict__ deg) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  atomicAdd(deg + src[e], 1.0f);
  atomicAdd(deg + dst[e], 1.0f);
}

// ====================================================================================
// layout conversion: row-major fp32 [rows,64] <-> tile images
// ====================================================================================
template <int HP>
__global__ void __launch_bounds__(256) tc_pack_state_kernel(const float* __restrict__ h, const float* __restrict__ c,
                                                            int64_t n_rows, int64_t n_rows_pad,
                                                            uint8_t* __restrict__ state) {
  // one thread per (row, pair of columns); padded rows are zero-filled
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows_pad * 32) return;
  const int64_t row = i >> 5;
  const int col = static_cast<int>(i & 31) * 2;
  uint8_t* tile = state + (row / TILE_ROWS) * tile_bytes(HP);
  const int r = static_cast<int>(row % TILE_ROWS);
  if (h) {
    float2 hv = make_float2(0.f, 0.f);
    if (row < n_rows) hv = *reinterpret_cast<const float2*>(h + row * D + col);
    uint32_t hi, lo;
    ptx::split_bf16x2(hv.x, hv.y, hi, lo);
    *reinterpret_cast<uint32_t*>(tile + plane_off(r, col)) = hi;
    if (HP == 2) *reinterpret_cast<uint32_t*>(tile + PLANE_BYTES + plane_off(r, col)) = lo;
  }
  if (c) {
    float2 cv = make_float2(0.f, 0.f);
    if (row < n_rows) cv = *reinterpret_cast<const float2*>(c + row * D + col);
    *reinterpret_cast<float2*>(tile + HP * PLANE_BYTES + ct_off(r, col)) = cv;
  }
}

// E0 = E_init_MLP([W, C]) written straight into the edge tile images, c = 0 (model.py:33-43,
// graphnn.py:134-139): same arithmetic as simt_edge_init_kernel, but thread = row of a tile, so the
// 16-byte chunk stores of a warp are contiguous (512 B) instead of one 256-byte row per thread, and
// the separate pack / zero passes disappear.  One CTA per tile.
template <int HP>
__global__ void __launch_bounds__(TILE_ROWS) tc_edge_init_kernel(const float* __restrict__ W, const float* __restrict__ C,
                                                                 const float* __restrict__ einit_blob, int64_t n_rows,
                                                                 uint8_t* __restrict__ state) {
  // The 2,824 parameters of E_init_MLP sit contiguously in the blob in EInit's member order; staged in
  // shared memory they are read with warp-uniform (broadcast) LDS instead of 2,704 distinct
  // constant-bank operands per thread, which thrash the constant cache.
  __shared__ __align__(16) float wbuf[sizeof(EInit) / 4];
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(EInit) / 4); i += TILE_ROWS) wbuf[i] = einit_blob[i];
  __syncthreads();
  const EInit& P = *reinterpret_cast<const EInit*>(wbuf);
  const int r = threadIdx.x;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * TILE_ROWS + r;
  uint8_t* tile = state + static_cast<int64_t>(blockIdx.x) * tile_bytes(HP);
  const bool valid = row < n_rows;
  const float in0 = valid ? W[row] : 0.f, in1 = valid ? C[row] : 0.f;
  float a1[8], a2[16], a3[32];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = fmaxf(fmaf(in1, P.w1[1][j], fmaf(in0, P.w1[0][j], P.b1[j])), 0.f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float s = P.b2[j];
#pragma unroll
    for (int k = 0; k < 8; ++k) s = fmaf(a1[k], P.w2[k][j], s);
    a2[j] = fmaxf(s, 0.f);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float s = P.b3[j];
#pragma unroll
    for (int k = 0; k < 16; ++k) s = fmaf(a2[k], P.w3[k][j], s);
    a3[j] = fmaxf(s, 0.f);
  }
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = P.b4[ch * 8 + q];
#pragma unroll
    for (int k = 0; k < 32; ++k) {       // k outer: the eight weights of a k are two LDS.128
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = fmaf(a3[k], P.w4[k][ch * 8 + q], o[q]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = valid ? o[q] : 0.f;   // padded rows are zero like tc_pack_state_kernel leaves them
    uint4 hi, lo;
    split8(o, hi, lo);
    *reinterpret_cast<uint4*>(tile + ch * 2048 + r * 16) = hi;
    if (HP == 2) *reinterpret_cast<uint4*>(tile + PLANE_BYTES + ch * 2048 + r * 16) = lo;
  }
  float4* cg = reinterpret_cast<float4*>(tile + HP * PLANE_BYTES) + r;
#pragma unroll
  for (int q = 0; q < 16; ++q) cg[q * 128] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// zero the c tile of every state tile (graphnn.py:137)
template <int HP>
__global__ void __launch_bounds__(256) tc_zero_c_kernel(int64_t n_rows_pad, uint8_t* __restrict__ state) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // float4 index
  if (i >= n_rows_pad * 16) return;
  const int64_t tile = i / (TILE_ROWS * 16);
  reinterpret_cast<float4*>(state + tile * tile_bytes(HP) + HP * PLANE_BYTES)[i % (TILE_ROWS * 16)] =
      make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int HP>
__global__ void __launch_bounds__(256) tc_unpack_state_kernel(const uint8_t* __restrict__ state, int64_t n_rows,
                                                              float* __restrict__ h, float* __restrict__ c) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows * 32) return;
  const int64_t row = i >> 5;
  const int col = static_cast<int>(i & 31) * 2;
  const uint8_t* tile = state + (row / TILE_ROWS) * tile_bytes(HP);
  const int r = static_cast<int>(row % TILE_ROWS);
  if (h) {
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(tile + plane_off(r, col));
    float x0 = __uint_as_float(hi << 16), x1 = __uint_as_float(hi & 0xFFFF0000u);
    if (HP == 2) {
      const uint32_t lo = *reinterpret_cast<const uint32_t*>(tile + PLANE_BYTES + plane_off(r, col));
      x0 += __uint_as_float(lo << 16);
      x1 += __uint_as_float(lo & 0xFFFF0000u);
    }
    __stcs(reinterpret_cast<float2*>(h + row * D + col), make_float2(x0, x1));   // streaming: the copy is not read back soon
  }
  if (c) __stcs(reinterpret_cast<float2*>(c + row * D + col),
                *reinterpret_cast<const float2*>(tile + HP * PLANE_BYTES + ct_off(r, col)));
}

}  // namespace tspgnn
