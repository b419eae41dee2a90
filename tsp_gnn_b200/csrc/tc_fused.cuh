// Fused message-passing timestep on CTA pairs (tcgen05 cta_group::2).
//
// One launch = one iteration of while_body (graphnn.py:142-173) for every edge and vertex row:
//
//   edge tile   : x = mV_in[src] + mV_in[dst]  ->  LayerNorm-LSTM cell (E.h, E.c in place)
//                 -> E_msg_V hidden layers on the FRESH h' (never re-read from global memory)
//                 -> scatter-add of the last hidden activations into xV_out      (folded output layer)
//   vertex tile : x = xV_in (read and cleared) ->  LayerNorm-LSTM cell (V.h, V.c in place)
//                 -> V_msg_E on the fresh h'   ->  mV_out
//
// Both kinds of tile only consume what the PREVIOUS launch produced (mV_in / xV_in are the other
// halves of two double buffers), so a timestep has one grid-wide dependency instead of the two of the
// K2 / K1 sequence, and E.h is read once.  What made this impossible on one SM is shared memory: the
// hi+lo bf16 images of the LSTM kernel (128 KB) and of three MLP layers (48 KB) leave no room for
// operands.  A CTA pair splits every B image across its two SMs (tcgen05.mma.cta_group::2, M = 256:
// each CTA supplies its own 128 rows of A and HALF of the output features of B and receives its
// 128 rows x N columns of D in its own TMEM): 88 KB of weights per CTA, 128 KB of operand slots.
//
// Warp roles per CTA (384 threads), both CTAs of the pair run the same program on their own tile
// (tile 2p + rank of tile pair p):
//   warp  0        : even CTA: issues every tcgen05.mma of the pair (event driven, f_mma); odd CTA: forwards
//                    its operand barriers (weights, h planes, x operand) to their twins in the even CTA
//   warps 1-3      : producers of the x operand (gather / read-and-clear)
//   warps 4-7, 8-11: two chain warpgroups, tile g -> warpgroup g & 1; thread = row = TMEM lane.
//                    cell epilogue (k1_cell_tile) -> MLP layer epilogues -> scatter / store
//   (the chain warps carry the critical path and get the higher warp ids: the scheduler prefers them, so
//   the gather fills issue slots instead of taking them)
// Shared-memory slots (32 KB each): slot 0 = x operand; slots 1.. = ring of h slots.  A tile's h
// slot is its home for the whole chain: h planes (bulk copy) -> A operand of the LSTM MMA -> new h'
// planes (written by the cell epilogue) -> hidden activations of every MLP layer, in place -> fp32
// staging of the scatter.  The warpgroup that retires the slot issues the bulk copy of the tile that
// uses it next.
// TMEM: 2 x 256 columns, half e belongs to warpgroup e: z of the LSTM, then the 64-column
// accumulators of the MLP layers in the (consumed) first columns.
#pragma once
#include "tc_kernels.cuh"

namespace tspgnn {

struct FArgs {
  uint8_t* stateE;
  uint8_t* stateV;
  const uint8_t* wlE;     // LSTM images per CTA rank: [rank][plane][kblock] x (128 features x 64 k bf16) = 16 KB each
  const uint8_t* wlV;     //   (vertex cell with the folded E_msg_V output layer)
  const uint8_t* wmE;     // MLP images per CTA rank: [rank][layer][plane] x (32 features x 64 k bf16) = 4 KB each
  const uint8_t* wmV;
  // message double buffers: timestep t reads half t & 1 and writes half (t + 1) & 1
  float* mVb[2];          // vertex messages [sumV][64]
  float* xVb[2];          // summed edge activations; read and cleared by vertex tiles
  const int32_t* src;
  const int32_t* dst;
  int64_t nE, nV;
  int pairsE, pairsV;     // tile pairs (the state images are allocated for 2 * pairs tiles)
  int e_clusters;         // clusters [0, e_clusters) own edge tile pairs, the others vertex tile pairs
  int clampE, clampV;
  int n_steps;            // timesteps of this launch (the kernel is persistent over them)
  int skip_last_mlp;      // nobody consumes the messages of the state after the last timestep
  int dbg;                // measurement aid (results are wrong when set): 1 = no reductions, 2 = no scatter / store,
                          // 4 = never any messages
  unsigned int* grid_ctr; // grid barrier between timesteps; zero at launch (cleared by the message kernel before)
  const float* vdeg;
  const float* ln_tab;
  const float* bias_tab;
  long long* timeline;
};

// Grid-wide barrier between two timesteps of the persistent kernel (every CTA is resident: the grid
// never exceeds the SM count).  Same pattern as cooperative groups' grid sync: CTA barrier, one thread
// publishes (release) and polls (acquire) a counter in global memory, CTA barrier.  The gpu-scope
// fences also invalidate this SM's L1, so the messages other SMs wrote are re-read from L2.
__device__ __forceinline__ void f_grid_sync(unsigned int* ctr, unsigned int target) {
  ptx::tcgen05_fence_before();
  asm volatile("bar.sync 0;" ::: "memory");
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
#ifdef TSPGNN_DEBUG_WAIT
      if (clock64() - t0 > 80000000LL) {
        ptx::wait_timeout_record(0xFFFFF, target);
        break;
      }
#else
      if (clock64() - t0 > 4000000000LL) __trap();      // a CTA that is not resident would hang the grid: fail loudly
#endif
    } while (seen < target);
    __threadfence();
  }
  asm volatile("bar.sync 0;" ::: "memory");
  ptx::tcgen05_fence_after();
}

// per-CTA geometry of the persistent loop
struct FGeo {
  int p0, rank, ntiles, nh, n_steps, nl_full, skip_last, nctas;
  bool boot1;     // edge CTAs with >= 2 tiles: the x operand of the timestep's SECOND tile is built by warpgroup 1 at the
                  // start of the timestep, in the h slot of the third tile (whose bulk copy waits until it is consumed)
  __device__ __forceinline__ int boot1_slot(int base) const { return 1 + (base + 2) % nh; }
  __device__ __forceinline__ int nl(int t) const { return (skip_last && t == n_steps - 1) ? 0 : nl_full; }
  __device__ __forceinline__ int tile(int n) const { return 2 * (p0 + n) + rank; }
};

template <int HP>
struct FSmem {
  static constexpr int WL_BYTES = HP * 2 * 16384;          // [plane][kblock] half images of the LSTM kernel
  static constexpr int WM_LAYER = HP * 4096;               // one MLP layer: [plane] half images
  static constexpr int WM_BYTES = 4 * WM_LAYER;
  static constexpr int SLOT_BYTES = 32768;                 // >= HP planes; = fp32 staging of 128 x 64 messages
  static constexpr int NSLOT = 4;
  static constexpr int NH_E = 3, NH_V = 2;                 // h slots in the ring (vertex CTAs keep slot 3 for tables)
  static constexpr int WL_OFF = 0;
  static constexpr int WM_OFF = WL_OFF + WL_BYTES;
  static constexpr int SLOT_OFF = WM_OFF + WM_BYTES;
  static constexpr int BAR_OFF = SLOT_OFF + NSLOT * SLOT_BYTES;
  static constexpr int NBAR = 18;
  static constexpr int TOTAL = BAR_OFF + 8 * NBAR + 16;
  static constexpr int DYN_BYTES = TOTAL + 128;
  // LayerNorm table (2560 B) + bias table (1024 B): edge CTAs use three MLP layers, so the fourth
  // layer's weight area is free; vertex CTAs use all four and put the tables into slot 3
  static constexpr int TAB_OFF_E = WM_OFF + 3 * WM_LAYER;
  static constexpr int TAB_OFF_V = SLOT_OFF + 3 * SLOT_BYTES;
  static constexpr int TAB_BYTES = 2 * 5 * D * 4 + 4 * D * 4;
};
static_assert(FSmem<2>::DYN_BYTES <= 232448, "fused kernel shared memory budget (227 KB)");
static_assert(FSmem<1>::TAB_BYTES <= FSmem<1>::WM_LAYER, "tables must fit the unused MLP layer area");

struct FBars {
  uint64_t* w;            // weight images landed
  uint64_t* x_full;       // x operand written (3 producer warps)
  uint64_t* x_empty;      // x operand consumed (tcgen05.commit)
  uint64_t* h_full;       // [3] h planes landed (bulk copy)
  uint64_t* acc_full;     // [2] an MMA group of warpgroup e completed (LSTM z, then every MLP layer)
  uint64_t* act_ready;    // [2] warpgroup e of BOTH CTAs: operand of the next MMA group written / accumulator drained
                          //     (only the even CTA's copy is used; the odd CTA's warps arrive on it remotely)
  uint64_t* boot;         // [2] x operand of the timestep's first tile (and, on edge CTAs, of its second tile, built
                          //     in the third h slot while that is still free) written by the chain warps
  // twins in the EVEN CTA, arrived by the odd CTA's relay warp
  uint64_t* p_w;
  uint64_t* p_x_full;
  uint64_t* p_h_full;     // [3]
  uint64_t* p_boot1;      // twin of boot[1]
};

__device__ __forceinline__ FBars f_bars(uint8_t* base) {
  uint64_t* b = reinterpret_cast<uint64_t*>(base);
  FBars r;
  r.w = b;
  r.x_full = b + 1;
  r.x_empty = b + 2;
  r.h_full = b + 3;
  r.acc_full = b + 6;
  r.act_ready = b + 8;
  r.boot = b + 16;
  r.p_w = b + 11;
  r.p_x_full = b + 12;
  r.p_h_full = b + 13;
  r.p_boot1 = b + 10;
  return r;
}

// ---- MMA issuer (even CTA) -----------------------------------------------------------------------
// Event driven: the two chain warpgroups are two independent streams of MMA work,
//     stream e:  [ layer 0 .. nl-1 of tile g,  LSTM(g + 2) ]  for the tiles g with g & 1 == e
// and the issuer serves whichever stream's next item is ready (non-blocking mbarrier tests), so a
// warpgroup in its latency-critical layer chain never queues behind the other one's items.  (A
// fixed issue order made the issuer the critical path: every layer step costs ~3 k cycles of
// epilogue + synchronisation latency, during which the other stream's requests waited.)  LSTM items are
// issued in tile order (the x operand slot is filled in tile order).
// Tiles are numbered g = t * ntiles + n over the whole launch: warpgroup, accumulator half, h slot and
// every barrier parity follow from g, so nothing is re-initialised between timesteps.
// act_ready[e] counts the four warps of warpgroup e of BOTH CTAs (the odd CTA's warps arrive remotely),
// the h / x operand barriers of the odd CTA are forwarded by its warp 8 (f_forward).
template <int HP>
__device__ __forceinline__ void f_mma(const FArgs& a, uint8_t* smem, const FBars& b, uint32_t tmem, const FGeo& geo) {
  using L = FSmem<HP>;
  constexpr uint32_t IDESC_LSTM = ptx::umma_idesc_bf16(256, 256);
  constexpr uint32_t IDESC_MLP = ptx::umma_idesc_bf16(256, 64);
  constexpr int NCOMB = (HP == 2) ? 3 : 1;
  const bool leader_lane = ptx::elect_one();
  const int nh = geo.nh, ntiles = geo.ntiles;
  const uint64_t slot_desc0 = ptx::umma_desc_k_nosw(ptx::smem_u32(smem + L::SLOT_OFF), 2048, 128);
  const uint64_t wl_desc0 = ptx::umma_desc_k_nosw(ptx::smem_u32(smem + L::WL_OFF), 2048, 128);
  const uint64_t wm_desc0 = ptx::umma_desc_k_nosw(ptx::smem_u32(smem + L::WM_OFF), 512, 128);
  long long* tl = nullptr;
  int gbase = 0;       // first tile number of the current timestep (for the trace only)

  // one k-block (h: half 0, x: half 1) of the LSTM contraction of tile g
  auto issue_lstm_half = [&](int g, int half) {
    const int e = g & 1, hs = g % nh;
    const uint32_t d_tmem = tmem + e * 256;
    const int kb = 1 - half;      // the h k-block first: its bulk copy lands long before the gathered x operand is built
    ptx::tcgen05_fence_after();
    if (ptx::elect_one()) {
      const int slot = (half == 0) ? 1 + hs : ((geo.boot1 && g == gbase + 1) ? geo.boot1_slot(gbase) : 0);
      const int pa_[3] = {1, 0, 0}, pb_[3] = {0, 1, 0};     // (A plane, B plane): cross terms first, then hi*hi
      const uint64_t aslot = slot_desc0 + static_cast<uint32_t>((slot * L::SLOT_BYTES) >> 4);
#pragma unroll
      for (int cb = 0; cb < NCOMB; ++cb) {
        const int pa = (HP == 2) ? pa_[cb] : 0, pb = (HP == 2) ? pb_[cb] : 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_bf16_ss_pair(d_tmem, aslot + ((pa * PLANE_BYTES + k * 4096) >> 4),
                                 wl_desc0 + (((pb * 2 + kb) * 16384 + k * 4096) >> 4), IDESC_LSTM,
                                 (half | cb | k) ? 1u : 0u);
      }
      if (half == 1) {
        ptx::umma_commit_pair(b.x_empty);
        ptx::umma_commit_pair(&b.acc_full[e]);
      }
    }
    __syncwarp();
    tl_mark(tl, 2, g - gbase, 1 + half);
  };
  auto h_ready = [&](int g) {
    const int hs = g % nh;
    const uint32_t par = (g / nh) & 1;
    return ptx::mbar_test(&b.h_full[hs], par) && ptx::mbar_test(&b.p_h_full[hs], par);
  };
  // x operand of tile g.  x_full is consumed strictly one phase at a time (xf counts them over the launch);
  // the second tile of a boot1 timestep is NOT on x_full: its operand is ready together with the first
  // tile's, and two completions in a row would run a whole phase ahead of this (parity-testing) consumer.
  uint32_t xf = 0;
  int tstep = 0;
  auto x_is_boot1 = [&](int g) { return geo.boot1 && g == gbase + 1; };
  auto x_ready = [&](int g) {
    if (x_is_boot1(g)) return ptx::mbar_test(&b.boot[1], tstep & 1) && ptx::mbar_test(b.p_boot1, tstep & 1);
    return ptx::mbar_test(b.x_full, xf & 1) && ptx::mbar_test(b.p_x_full, xf & 1);
  };

  ptx::mbar_wait(b.w, 0);
  ptx::mbar_wait(b.p_w, 0);
  uint32_t ar[2] = {0, 0};      // act_ready phases consumed per stream, over the whole launch
  for (int t = 0; t < geo.n_steps; ++t) {
    const int nl = (a.dbg & 4) ? 0 : geo.nl(t);
    const int base = t * ntiles, end = base + ntiles;
    gbase = base;
    tstep = t;
    tl = (t == 0 && leader_lane) ? a.timeline : nullptr;
    // per stream: tile, next layer (nl = the LSTM item of tile + 2 comes next)
    int tile_[2], layer_[2] = {0, 0};
    bool active[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      tile_[e] = base + ((e ^ base) & 1);
      active[e] = tile_[e] < end;
    }
    const bool had[2] = {active[0], active[1]};
    // LSTM items strictly in tile order: lstm_g = next tile whose LSTM has to be issued, lstm_h = its h
    // k-block is already issued; the first two tiles of a timestep need no drained accumulator
    int lstm_g = base;
    bool lstm_h = false;
    int drained_ok = (ntiles > 1) ? base + 2 : base + 1;      // LSTM items below this tile number may be issued
    long long spin0 = clock64();
    int nap = 0;
    while (active[0] || active[1] || lstm_g < drained_ok) {
      bool progress = false;
      // ---- the LSTM item at the head of the tile order ----
      if (lstm_g < drained_ok) {
        if (!lstm_h) {
          if (h_ready(lstm_g)) {
            tl_mark(tl, 2, lstm_g - gbase, 0);
            issue_lstm_half(lstm_g, 0);
            lstm_h = true;
            progress = true;
          }
        } else if (x_ready(lstm_g)) {
          if (!x_is_boot1(lstm_g)) ++xf;
          issue_lstm_half(lstm_g, 1);
          lstm_h = false;
          ++lstm_g;
          progress = true;
        }
      }
      // ---- the layer chains ----
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (!active[e]) continue;
        const int g = tile_[e];
        if (layer_[e] < nl) {
          if (!ptx::mbar_test(&b.act_ready[e], ar[e] & 1)) continue;
          ++ar[e];
          const int l = layer_[e], hs = g % nh;
          ptx::tcgen05_fence_after();
          if (ptx::elect_one()) {
            const uint64_t adesc = slot_desc0 + static_cast<uint32_t>(((1 + hs) * L::SLOT_BYTES) >> 4);
            const uint64_t bdesc = wm_desc0 + static_cast<uint32_t>((l * L::WM_LAYER) >> 4);
            const int pa_[3] = {1, 0, 0}, pb_[3] = {0, 1, 0};
#pragma unroll
            for (int cb = 0; cb < NCOMB; ++cb) {
              const int pa = (HP == 2) ? pa_[cb] : 0, pb = (HP == 2) ? pb_[cb] : 0;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_bf16_ss_pair(tmem + e * 256, adesc + ((pa * PLANE_BYTES + k * 4096) >> 4),
                                       bdesc + ((pb * 4096 + k * 1024) >> 4), IDESC_MLP, (cb | k) ? 1u : 0u);
            }
            ptx::umma_commit_pair(&b.acc_full[e]);
          }
          __syncwarp();
          if (l < 4) tl_mark(tl, 2, g - gbase, 3 + l);
          ++layer_[e];
          progress = true;
        } else {
          // the chain of tile g is issued; its accumulator half is reusable once the warpgroup has drained it
          if (g + 2 >= end) {
            active[e] = false;        // (the stream's last drained arrival of the timestep is consumed below)
            progress = true;
            continue;
          }
          if (drained_ok != g + 2) continue;      // the other stream's LSTM item comes first in tile order
          if (!ptx::mbar_test(&b.act_ready[e], ar[e] & 1)) continue;
          ++ar[e];
          drained_ok = g + 3;
          tile_[e] = g + 2;
          layer_[e] = 0;
          progress = true;
        }
      }
      if (progress) {
        spin0 = clock64();
      } else {
        // nothing ready: sleep on a layer-chain barrier instead of spinning (this warp shares its
        // scheduler with a warp of each chain warpgroup), alternating between the two streams
        nap ^= 1;
        const int e = (active[nap] && layer_[nap] < nl) ? nap : nap ^ 1;
        if (active[e] && layer_[e] < nl) ptx::mbar_try_wait_ns(&b.act_ready[e], ar[e] & 1, a.dbg >> 8 ? (a.dbg >> 8) : 160);
        else __nanosleep(a.dbg >> 8 ? (a.dbg >> 8) : 100);
#ifdef TSPGNN_DEBUG_WAIT
        if (clock64() - spin0 > 40000000LL) {
          // issuer stuck: record (pending LSTM tile - base, its h issued?, drained_ok - base, per stream tile - base / layer)
          ptx::wait_timeout_record(0xE0000u | ((lstm_g - base) << 12) | (lstm_h ? 0x800 : 0) | ((drained_ok - base) << 6) |
                                       ((tile_[0] - base) << 3) | (tile_[1] - base),
                                   (layer_[0] << 4 | layer_[1]) & 1);
          break;
        }
#else
        if (clock64() - spin0 > 4000000000LL) __trap();      // protocol bug: fail loudly instead of hanging
#endif
      }
    }
    // the last tile of each stream: its drained arrival closes the stream's phase count for this timestep
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (had[e]) {
        ptx::mbar_wait(&b.act_ready[e], ar[e] & 1);
        ++ar[e];
      }
    if (t + 1 < geo.n_steps) f_grid_sync(a.grid_ctr, static_cast<unsigned int>(t + 1) * geo.nctas);
  }
}

// ---- odd CTA, warp 8: forwards its operand barriers to the twins in the even CTA -----------------
__device__ __forceinline__ void f_forward(const FArgs& a, const FBars& b, const FGeo& geo) {
  auto forward = [&](uint64_t* local, uint64_t* twin, uint32_t parity) {
    ptx::mbar_wait(local, parity);
    if (ptx::elect_one()) ptx::mbar_arrive_remote(ptx::mapa_u32(ptx::smem_u32(twin), 0));
    __syncwarp();
  };
  forward(b.w, b.p_w, 0);
  uint32_t xf = 0;      // x_full phases forwarded (see f_mma::x_ready)
  for (int t = 0; t < geo.n_steps; ++t) {
    for (int n = 0; n < geo.ntiles; ++n) {
      const int g = t * geo.ntiles + n, hs = g % geo.nh;
      forward(&b.h_full[hs], &b.p_h_full[hs], (g / geo.nh) & 1);
      if (geo.boot1 && n == 1) {
        forward(&b.boot[1], b.p_boot1, t & 1);
      } else {
        forward(b.x_full, b.p_x_full, xf & 1);
        ++xf;
      }
    }
    if (t + 1 < geo.n_steps) f_grid_sync(a.grid_ctr, static_cast<unsigned int>(t + 1) * geo.nctas);
  }
}

// ---- producers of the x operand (single slot) --------------------------------------------------
template <int HP, bool IS_V>
__device__ __forceinline__ void f_producer(const FArgs& a, uint8_t* smem, const FBars& b, const FGeo& geo, int gw,
                                           int lane) {
  using L = FSmem<HP>;
  uint8_t* xslot = smem + L::SLOT_OFF;
  uint8_t* state = IS_V ? a.stateV : a.stateE;
  const int r8 = lane & 7;
  const int ntiles = geo.ntiles;
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
  for (int t = 0; t < geo.n_steps; ++t) {
    long long* tl = (t == 0 && gw == 0 && lane == 0) ? a.timeline : nullptr;
    const float* mV_in = a.mVb[t & 1];
    float* xV_in = a.xVb[t & 1];
    const int base = t * ntiles;
    auto load_h = [&](int n) {
      const int hs = (base + n) % geo.nh;
      ptx::mbar_arrive_expect_tx(&b.h_full[hs], HP * PLANE_BYTES);
      ptx::bulk_g2s(smem + L::SLOT_OFF + (1 + hs) * L::SLOT_BYTES,
                    state + static_cast<int64_t>(geo.tile(n)) * tile_bytes(HP), HP * PLANE_BYTES, &b.h_full[hs]);
    };
    if (gw == 0 && lane == 0) {
      // h planes of the first tiles of the timestep (every slot is free: the previous timestep is
      // complete); later ones are fetched by the warpgroup that retires a slot.  With boot1 the third
      // tile's slot first holds the x operand of the second tile: its copy is issued further down.
      for (int n = 0; n < ntiles && n < geo.nh; ++n)
        if (!(geo.boot1 && n == 2)) load_h(n);
    }
    __syncwarp();
    const int first_gather = geo.boot1 ? 2 : 1;      // earlier tiles: x operand built by the chain warps
    int si[6], di[6];
#pragma unroll
    for (int gi = 0; gi < 6; ++gi) si[gi] = di[gi] = 0;
    if (ntiles > first_gather) load_idx(geo.tile(first_gather), si, di);
    for (int n = 0; n < ntiles; ++n) {
      if (geo.boot1 && n == 1) continue;      // operand and its barrier (boot[1]) come from warpgroup 1
      const int tile = geo.tile(n), g = base + n;
      int sn[6], dn[6];
#pragma unroll
      for (int gi = 0; gi < 6; ++gi) sn[gi] = dn[gi] = 0;
      if (n >= first_gather && n + 1 < ntiles) load_idx(tile + 2, sn, dn);   // next tile's column indices, a tile ahead
      tl_mark(tl, 3, n, 0);
      if (n >= first_gather) {
        // phase g - 1 of x_empty = LSTM(g - 1) has consumed its x operand.  Parity waits must be taken in
        // order: the first gather of a boot1 timestep also passes the phase of the tile it skipped
        if (geo.boot1 && n == 2) ptx::mbar_wait(b.x_empty, (g - 2) & 1);
        ptx::mbar_wait(b.x_empty, (g - 1) & 1);
      }
      tl_mark(tl, 3, n, 1);
      if (geo.boot1 && n == 2 && gw == 0 && lane == 0) load_h(2);      // LSTM(1) has consumed the x operand parked there
      if (n < first_gather) {
        ptx::mbar_wait(&b.boot[0], t & 1);      // (n == 0) every producer warp waits: each one arrives on x_full below
      } else {
        k1_fill_x<HP, IS_V, false>(xslot, gw, lane, static_cast<int64_t>(tile) * TILE_ROWS, mV_in, xV_in, si, di);
        ptx::fence_proxy_async_smem();
#pragma unroll
        for (int gi = 0; gi < 6; ++gi) {
          si[gi] = sn[gi];
          di[gi] = dn[gi];
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(b.x_full);
      tl_mark(tl, 3, n, 2);
    }
    if (t + 1 < geo.n_steps) f_grid_sync(a.grid_ctr, static_cast<unsigned int>(t + 1) * geo.nctas);
  }
}

// ---- chain warpgroups -----------------------------------------------------------------------------
// IS_V: false = edge tiles (three hidden layers, scatter-add of a3), true = vertex tiles (four layers,
// message store).
template <int HP, bool IS_V, bool CLAMP>
__device__ __forceinline__ void f_chain(const FArgs& a, uint8_t* smem, const FBars& b, uint32_t tmem, const FGeo& geo,
                                        int warp, int lane, uint32_t ln_s, uint32_t bias_s) {
  using L = FSmem<HP>;
  constexpr int NL = IS_V ? 4 : 3;
  constexpr int NH = IS_V ? L::NH_V : L::NH_E;
  const int e = (warp >> 2) - 1, q4 = warp & 3;
  const int r = q4 * 32 + lane;
  const int ntiles = geo.ntiles;
  const uint32_t t_acc = tmem + (static_cast<uint32_t>(q4 * 32) << 16) + e * 256;
  uint8_t* state = IS_V ? a.stateV : a.stateE;
  const int64_t n_rows = IS_V ? a.nV : a.nE;
  const uint32_t slots_s = ptx::smem_u32(smem + L::SLOT_OFF);
  // the issuer lives in the even CTA: every warp of this warpgroup of BOTH CTAs arrives on its act_ready[e]
  const uint32_t act_ready_leader = ptx::mapa_u32(ptx::smem_u32(&b.act_ready[e]), 0);
  uint32_t af = 0;              // acc_full phases consumed by this warpgroup, over the whole launch
  for (int t = 0; t < geo.n_steps; ++t) {
    const int nl = (a.dbg & 4) ? 0 : geo.nl(t);
    long long* tl = (t == 0 && q4 == 0 && lane == 0) ? a.timeline : nullptr;
    const float* mV_in = a.mVb[t & 1];
    float* mV_out = a.mVb[(t + 1) & 1];
    float* xV_in = a.xVb[t & 1];
    float* xV_out = a.xVb[(t + 1) & 1];
    const int base = t * ntiles;
    if (ntiles > 0) {
      // x operands of the timestep's first tile(s): built by the chain warps, which have nothing else to
      // do until the first accumulator is ready -- warpgroup 0 the first tile and warpgroup 1 the second
      // one (edge CTAs), or all eight warps the first tile
      if (geo.boot1) {
        if (e == 0)
          k1_boot_fill_tile<HP, IS_V, false, 4>(mV_in, xV_in, a.src, a.dst, smem + L::SLOT_OFF, q4, lane, geo.tile(0));
        else
          k1_boot_fill_tile<HP, IS_V, false, 4>(mV_in, xV_in, a.src, a.dst,
                                                smem + L::SLOT_OFF + geo.boot1_slot(base) * L::SLOT_BYTES, q4, lane,
                                                geo.tile(1));
      } else {
        k1_boot_fill_tile<HP, IS_V, false, 8>(mV_in, xV_in, a.src, a.dst, smem + L::SLOT_OFF, warp - 4, lane, geo.tile(0));
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&b.boot[geo.boot1 ? e : 0]);
    }
    for (int n = (e ^ base) & 1; n < ntiles; n += 2) {
      const int tile = geo.tile(n);
      const int64_t row0 = static_cast<int64_t>(tile) * TILE_ROWS;
      const int hs = (base + n) % NH;
      const uint32_t b_s = slots_s + (1 + hs) * L::SLOT_BYTES;
      uint8_t* gtile = state + static_cast<int64_t>(tile) * tile_bytes(HP);
      // endpoints of this warp's 32 rows, fetched a whole chain ahead of the scatter that uses them
      int my_s = -1, my_d = -1;
      if (!IS_V && nl > 0 && row0 + r < n_rows) {
        my_s = __ldg(a.src + row0 + r);
        my_d = __ldg(a.dst + row0 + r);
      }
      const float* vdeg_row = (IS_V && a.vdeg != nullptr) ? a.vdeg + row0 + r : nullptr;
      k1_cell_tile<HP, IS_V ? 0 : 1, CLAMP, true>(gtile, r, lane, t_acc, &b.acc_full[e], af & 1, nullptr, ln_s, vdeg_row,
                                                  b_s, tl, e, n);
      ++af;
      // h' planes are in the slot (first MLP operand) / the accumulator is drained; they are also in
      // global memory for the next timestep's bulk copy, an async-proxy read: hence the all-space proxy
      // fence below
      ptx::fence_proxy_async_smem();
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_remote_relaxed(act_ready_leader);
      ptx::fence_proxy_async_all();      // (global h' planes; off the critical path: the warpgroup waits for the MMA now)
      if (nl > 0) {
        float v[64];
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
          ptx::mbar_wait(&b.acc_full[e], af & 1);
          ++af;
          if (e == 0 && l < 3) tl_mark(tl, 3, n, 3 + l);      // (trace: accumulator of layer l seen by warpgroup 0)
          ptx::tcgen05_fence_after();
          ptx::tmem_ld64(t_acc, v);
          const uint32_t bl = bias_s + l * 256;       // broadcast LDS.128: four bias values per load
          if (l < NL - 1) {
            // hidden layer feeding the next MMA: bias + ReLU + bf16 split, in place (this layer's MMA has
            // finished reading the slot: its completion is what acc_full signals)
            const uint32_t nxt = b_s + r * 16;
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
                ptx::split_bf16x2_p(x, hi[p], lo[p]);
              }
              ptx::sts128(nxt + ch * 2048, make_uint4(hi[0], hi[1], hi[2], hi[3]));
              if (HP == 2) ptx::sts128(nxt + PLANE_BYTES + ch * 2048, make_uint4(lo[0], lo[1], lo[2], lo[3]));
            }
            ptx::fence_proxy_async_smem();
          } else {
            // last layer of the chain, kept in fp32: edge tiles bias + ReLU (a3 of the folded MLP),
            // vertex tiles bias only (linear output layer, graphnn.py:17)
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float4 bb = ptx::lds128f(bl + q * 16);
              float2 x0 = __fadd2_rn(make_float2(v[4 * q], v[4 * q + 1]), make_float2(bb.x, bb.y));
              float2 x1 = __fadd2_rn(make_float2(v[4 * q + 2], v[4 * q + 3]), make_float2(bb.z, bb.w));
              if (!IS_V) {
                x0 = ptx::relu2(x0);
                x1 = ptx::relu2(x1);
              }
              v[4 * q] = x0.x; v[4 * q + 1] = x0.y; v[4 * q + 2] = x1.x; v[4 * q + 3] = x1.y;
            }
          }
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote_relaxed(act_ready_leader);   // next operand written / accumulator drained
          tl_mark(tl, e, n, 6);
        }
        // stage the fp32 rows in the slot (its last MMA has completed), then every warp walks its own
        // 32 rows, two rows per instruction (a half-warp covers the 256 bytes of a row)
        if (!(a.dbg & 2)) {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            ptx::sts128f(b_s + stage_off(r, q), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
          __syncwarp();
          const int64_t g0 = row0 + q4 * 32;
          const int hw = lane >> 4, c16 = lane & 15;
          if (IS_V) {
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
              const int rr = 2 * i + hw;
              if (g0 + rr < n_rows) {
                const float4 m = ptx::lds128f(b_s + stage_off(q4 * 32 + rr, c16));
                *reinterpret_cast<float4*>(mV_out + (g0 + rr) * D + 4 * c16) = m;
              }
            }
          } else {
            // dst side: one vector reduction per row; src side accumulated over runs of equal src
            // (rows of a complete graph are sorted by src, instance_loader.py:60)
            const bool red = !(a.dbg & 1);
            int cur_s = -1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
              const int rr = 2 * i + hw;
              const int s = __shfl_sync(0xffffffffu, my_s, rr);
              const int d = __shfl_sync(0xffffffffu, my_d, rr);
              if (s >= 0) {
                const float4 m = ptx::lds128f(b_s + stage_off(q4 * 32 + rr, c16));
                if (red) ptx::red_add_v4(xV_out + static_cast<int64_t>(d) * D + 4 * c16, m);
                if (s != cur_s) {
                  if (cur_s >= 0 && red) ptx::red_add_v4(xV_out + static_cast<int64_t>(cur_s) * D + 4 * c16, acc);
                  cur_s = s;
                  acc = m;
                } else {
                  acc.x += m.x; acc.y += m.y; acc.z += m.z; acc.w += m.w;
                }
              }
            }
            if (cur_s >= 0 && red) ptx::red_add_v4(xV_out + static_cast<int64_t>(cur_s) * D + 4 * c16, acc);
          }
        }
      }
      // every warp of the warpgroup is done with the slot: it becomes the home of tile n + NH
      asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");
      if (q4 == 0 && lane == 0 && n + NH < ntiles) {
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive_expect_tx(&b.h_full[hs], HP * PLANE_BYTES);
        ptx::bulk_g2s(smem + L::SLOT_OFF + (1 + hs) * L::SLOT_BYTES,
                      state + static_cast<int64_t>(geo.tile(n + NH)) * tile_bytes(HP), HP * PLANE_BYTES, &b.h_full[hs]);
      }
      tl_mark(tl, e, n, 7);
    }
    if (t + 1 < geo.n_steps) f_grid_sync(a.grid_ctr, static_cast<unsigned int>(t + 1) * geo.nctas);
  }
}

template <int HP>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_step_kernel(const FArgs a) {
  using L = FSmem<HP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const FBars b = f_bars(smem + L::BAR_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::BAR_OFF + 8 * L::NBAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cluster = static_cast<int>(blockIdx.x) >> 1, nclusters = static_cast<int>(gridDim.x) >> 1;
  const bool is_v = cluster >= a.e_clusters;
  FGeo geo;
  geo.rank = static_cast<int>(ptx::cluster_ctarank());
  int p1;
  if (is_v) tile_range(cluster - a.e_clusters, nclusters - a.e_clusters, a.pairsV, geo.p0, p1);
  else tile_range(cluster, a.e_clusters, a.pairsE, geo.p0, p1);
  geo.ntiles = p1 - geo.p0;               // this CTA's tiles: 2 * (p0 + n) + rank
  geo.nh = is_v ? L::NH_V : L::NH_E;
  geo.n_steps = a.n_steps;
  geo.nl_full = is_v ? 4 : 3;
  geo.skip_last = a.skip_last_mlp;
  geo.nctas = static_cast<int>(gridDim.x);
  geo.boot1 = !is_v && geo.ntiles > 1;
  const int tab_off = is_v ? L::TAB_OFF_V : L::TAB_OFF_E;

  if (tid == 0) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    for (int i = 0; i < L::NBAR; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::mbar_init(b.x_full, NUM_GATHER_WARPS);
    ptx::mbar_init(&b.act_ready[0], 8);      // four warps of the warpgroup in each CTA of the pair
    ptx::mbar_init(&b.act_ready[1], 8);
    ptx::mbar_init(&b.boot[0], geo.boot1 ? 4 : 8);
    ptx::mbar_init(&b.boot[1], 4);
    ptx::fence_mbar_init();
    // weight images: parameters, not produced by the preceding kernel
    const int wm_bytes = (is_v ? 4 : 3) * L::WM_LAYER;
    const uint8_t* wl = (is_v ? a.wlV : a.wlE) + static_cast<int64_t>(geo.rank) * L::WL_BYTES;
    const uint8_t* wm = (is_v ? a.wmV : a.wmE) + static_cast<int64_t>(geo.rank) * L::WM_BYTES;
    ptx::mbar_arrive_expect_tx(b.w, L::WL_BYTES + wm_bytes);
    for (int off = 0; off < L::WL_BYTES; off += 32768) ptx::bulk_g2s(smem + L::WL_OFF + off, wl + off, 32768, b.w);
    ptx::bulk_g2s(smem + L::WM_OFF, wm, wm_bytes, b.w);
  }
  if (warp == 0) ptx::tmem_alloc_pair(tmem_slot, 512);
  {
    // LayerNorm parameters of this CTA's cell as the epilogue wants them (see tc_lnlstm_kernel) and
    // the biases of its message MLP
    float* tab_sm = reinterpret_cast<float*>(smem + tab_off);
    const float* ln = a.ln_tab + (is_v ? 0 : 2 * 5 * D);
    const float* bias = a.bias_tab + (is_v ? 0 : 1) * 4 * D;
    for (int i = tid; i < 2 * 5 * D; i += TC_THREADS) tab_sm[i] = ln[i];
    for (int i = tid; i < 4 * D; i += TC_THREADS) tab_sm[2 * 5 * D + i] = bias[i];
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  // the barriers of BOTH CTAs are initialised and both TMEM allocations done before any remote
  // arrive or multicast commit can reach them
  ptx::cluster_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t ln_s = ptx::smem_u32(smem + tab_off);
  const uint32_t bias_s = ln_s + 2 * 5 * D * 4;
  // everything above overlapped the tail of the previous kernel (programmatic dependent launch);
  // from here on the recurrent state and the messages it produced are read
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();

  if (warp >= 4) {
    ptx::setmaxnreg_inc<200>();
    const bool clamp = (is_v ? a.clampV : a.clampE) != 0;
    if (is_v) {
      if (clamp) f_chain<HP, true, true>(a, smem, b, tmem, geo, warp, lane, ln_s, bias_s);
      else f_chain<HP, true, false>(a, smem, b, tmem, geo, warp, lane, ln_s, bias_s);
    } else {
      if (clamp) f_chain<HP, false, true>(a, smem, b, tmem, geo, warp, lane, ln_s, bias_s);
      else f_chain<HP, false, false>(a, smem, b, tmem, geo, warp, lane, ln_s, bias_s);
    }
  } else {
    ptx::setmaxnreg_dec<104>();
    if (warp == 0) {
      if (geo.rank == 0) f_mma<HP>(a, smem, b, tmem, geo);
      else f_forward(a, b, geo);
    } else {
      if (is_v) f_producer<HP, true>(a, smem, b, geo, warp - 1, lane);
      else f_producer<HP, false>(a, smem, b, geo, warp - 1, lane);
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  // the peer's shared memory and TMEM must stay alive until every MMA of the pair has completed
  ptx::cluster_sync();
  if (warp == 0) ptx::tmem_dealloc_pair(tmem, 512);
}

}  // namespace tspgnn
