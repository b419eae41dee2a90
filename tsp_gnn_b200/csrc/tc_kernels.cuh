// Tensor-core (tcgen05) kernels of the message-passing timestep.
//
//   K1  tc_lnlstm_kernel : gather/segment input -> z = [x,h].K on tcgen05 (TMEM accumulator,
//                          256 columns) -> 5x LayerNorm + gates in the epilogue -> new (c,h)
//                          (graphnn.py:155-170 for both variables of model.py:74-92)
//   K2  tc_mlp_kernel    : 4-layer message MLP chained through TMEM (graphnn.py:152-154), then
//                          E rows: scatter-add into xV (= EV^T . msg, model.py:76-83)
//                          V rows: store the vertex message consumed by K1's gather
//                          vote  : E_vote MLP, 64->1 tail in registers (model.py:107-128)
//
// HBM layout ("tile images"): recurrent state is stored per 128-row tile exactly as the
// bytes the kernels want in shared memory, so one bulk async copy (UBLKCP) stages a tile:
//   [h hi : 128 x 64 bf16, 128-B rows, 16-B chunks XOR-swizzled by (row & 7)]   16 KB
//   [h lo : same, bf16(h - hi)]                      (HP == 2 only)            16 KB
//   [c    : 128 x 64 fp32, 256-B rows, chunk ^= (row & 7)]                      32 KB
// The h planes are directly the K-major SWIZZLE_128B UMMA A operand.  HP = number of bf16
// planes: 2 -> every product is hi*hi + hi*lo + lo*hi (fp32-parity mode), 1 -> single bf16.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace tspgnn {

constexpr int IMG16_BYTES = TILE_ROWS * 128;   // one bf16 plane of a tile
constexpr int IMG32_BYTES = TILE_ROWS * 256;   // fp32 tile
constexpr int STAGE_LD = 66;                   // fp32 staging row stride (floats)
constexpr int STAGE_BYTES = 34 * 1024;         // >= 128*66*4

__host__ __device__ constexpr int tile_bytes(int hp) { return hp * IMG16_BYTES + IMG32_BYTES; }

// byte offset of element (row, col) inside a bf16 plane / fp32 tile image
__host__ __device__ inline uint32_t img16_off(int r, int col) {
  return static_cast<uint32_t>(r * 128 + (((col >> 3) ^ (r & 7)) << 4) + (col & 7) * 2);
}
__host__ __device__ inline uint32_t img32_off(int r, int col) {
  return static_cast<uint32_t>(r * 256 + (((col >> 2) ^ (r & 7)) << 4) + (col & 3) * 4);
}

__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fast_rcp(1.0f + __expf(-x)); }

struct K1Args {
  uint8_t* stateE;
  uint8_t* stateV;
  const uint8_t* wE;   // LSTM weight images, [plane][kblock] x (256 x 64 bf16) = 32 KB each
  const uint8_t* wV;
  const float* mV;     // vertex messages [sumV][64] (E rows gather two of them)
  float* xV;           // summed edge messages [sumV_pad][64]; V rows read and then zero it
  const int32_t* src;
  const int32_t* dst;
  int64_t nE, nV;
  int tilesE, tilesV, e_ctas;
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
};

// contiguous, balanced range of tiles for CTA `i` of `n`
__device__ __forceinline__ void tile_range(int i, int n, int tiles, int& t0, int& t1) {
  t0 = static_cast<int>((static_cast<int64_t>(i) * tiles) / n);
  t1 = static_cast<int>((static_cast<int64_t>(i + 1) * tiles) / n);
}

// ====================================================================================
// K1: LayerNorm-LSTM step
// ====================================================================================
template <int HP>
struct K1Smem {
  static constexpr int W_BYTES = HP * 2 * 256 * 128;       // [plane][kblock] 32 KB images
  static constexpr int X_OFF = W_BYTES;                    // x planes   (A operand, k-block 0)
  static constexpr int H_OFF = X_OFF + HP * IMG16_BYTES;   // h planes   (A operand, k-block 1)
  static constexpr int C_OFF = H_OFF + HP * IMG16_BYTES;   // c tile (contiguous after h: one tile image)
  static constexpr int BAR_OFF = C_OFF + IMG32_BYTES;
  static constexpr int TOTAL = BAR_OFF + 64;
  static constexpr int DYN_BYTES = TOTAL + 1024;           // slack for 1024-B alignment
};

template <int HP>
__global__ void __launch_bounds__(128, 1) tc_lnlstm_kernel(const K1Args a) {
  using L = K1Smem<HP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* xbuf = smem + L::X_OFF;
  uint8_t* hbuf = smem + L::H_OFF;
  uint8_t* cbuf = smem + L::C_OFF;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* bar_ld = bar_w + 1;
  uint64_t* bar_mma = bar_w + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_v = static_cast<int>(blockIdx.x) >= a.e_ctas;
  const int cell = is_v ? 0 : 1;
  int t0, t1;
  if (is_v) tile_range(blockIdx.x - a.e_ctas, gridDim.x - a.e_ctas, a.tilesV, t0, t1);
  else tile_range(blockIdx.x, a.e_ctas, a.tilesE, t0, t1);
  uint8_t* state = is_v ? a.stateV : a.stateE;
  const int64_t n_rows = is_v ? a.nV : a.nE;

  if (tid == 0) {
    ptx::mbar_init(bar_w, 1);
    ptx::mbar_init(bar_ld, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 256);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0 && t0 < t1) {
    const uint8_t* wimg = is_v ? a.wV : a.wE;
    ptx::mbar_arrive_expect_tx(bar_w, L::W_BYTES);
    for (int off = 0; off < L::W_BYTES; off += 32768) ptx::bulk_g2s(wsm + off, wimg + off, 32768, bar_w);
  }

  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(128, 256);
  const CellLN& ln = c_ln[cell];
  const int r = tid;                                  // row of the tile owned in the epilogue
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t phase = 0;

  for (int tile = t0; tile < t1; ++tile) {
    uint8_t* gtile = state + static_cast<int64_t>(tile) * tile_bytes(HP);
    const int64_t row0 = static_cast<int64_t>(tile) * TILE_ROWS;
    // ---- stage h planes + c: one contiguous tile image ------------------------------
    if (tid == 0) {
      ptx::mbar_arrive_expect_tx(bar_ld, tile_bytes(HP));
      ptx::bulk_g2s(hbuf, gtile, tile_bytes(HP), bar_ld);
    }
    // ---- build x planes: E rows x = mV[src]+mV[dst] (= EV.msg), V rows x = xV (then cleared)
    {
      int64_t grow = row0 + warp * 32 + lane;
      int my_s = 0, my_d = 0;
      if (!is_v && grow < n_rows) {
        my_s = a.src[grow];
        my_d = a.dst[grow];
      }
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const int row = warp * 32 + rr;
        const int64_t g = row0 + row;
        float2 x = make_float2(0.f, 0.f);
        if (is_v) {
          if (g < n_rows) {
            float2* p = reinterpret_cast<float2*>(a.xV + g * D) + lane;
            x = *p;
            *p = make_float2(0.f, 0.f);
          }
        } else {
          const int s = __shfl_sync(0xffffffffu, my_s, rr);
          const int d = __shfl_sync(0xffffffffu, my_d, rr);
          if (g < n_rows) {
            const float2 u = __ldg(reinterpret_cast<const float2*>(a.mV + static_cast<int64_t>(s) * D) + lane);
            const float2 w = __ldg(reinterpret_cast<const float2*>(a.mV + static_cast<int64_t>(d) * D) + lane);
            x = make_float2(u.x + w.x, u.y + w.y);
          }
        }
        uint32_t hi, lo;
        ptx::split_bf16x2(x.x, x.y, hi, lo);
        const uint32_t off = img16_off(row, 2 * lane);
        *reinterpret_cast<uint32_t*>(xbuf + off) = hi;
        if (HP == 2) *reinterpret_cast<uint32_t*>(xbuf + IMG16_BYTES + off) = lo;
      }
    }
    ptx::fence_proxy_async_smem();
    __syncthreads();
    // ---- MMA: z[128 x 256] = [x,h] . K ------------------------------------------------
    if (tid == 0) {
      if (tile == t0) ptx::mbar_wait(bar_w, 0);
      ptx::mbar_wait(bar_ld, phase);
      ptx::tcgen05_fence_after();
      uint32_t acc = 0;
      // (A plane, B plane): small cross terms first, then hi*hi
      constexpr int NCOMB = (HP == 2) ? 3 : 1;
      const int pa_[3] = {1, 0, 0}, pb_[3] = {0, 1, 0};
#pragma unroll
      for (int cb = 0; cb < NCOMB; ++cb) {
        const int pa = (HP == 2) ? pa_[cb] : 0, pb = (HP == 2) ? pb_[cb] : 0;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t abase = ptx::smem_u32((kb == 0 ? xbuf : hbuf) + pa * IMG16_BYTES);
          const uint32_t bbase = ptx::smem_u32(wsm + (pb * 2 + kb) * 32768);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ptx::umma_bf16_ss(tmem, ptx::umma_desc_k_sw128(abase + k * 32), ptx::umma_desc_k_sw128(bbase + k * 32),
                              IDESC, acc);
            acc = 1;
          }
        }
      }
      ptx::umma_commit(bar_mma);
    }
    __syncwarp();
    ptx::mbar_wait(bar_ld, phase);     // c tile visible to every thread
    ptx::mbar_wait(bar_mma, phase);
    ptx::tcgen05_fence_after();

    // ---- epilogue: thread r owns row r (TMEM lane r) ---------------------------------
    float mu[4], rs[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[64];
      ptx::tmem_ld64(t_lane + g * 64, v);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) s += v[j];
      const float m = s * (1.0f / 64);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const float t = v[j] - m;
        q = fmaf(t, t, q);
      }
      mu[g] = m;
      rs[g] = rsqrtf(q * (1.0f / 64) + LN_EPS);
    }
    float cn[64];
    uint8_t* crow = cbuf + r * 256;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      float vi[16], vj[16], vf[16], cv[16];
      ptx::tmem_ld16(t_lane + 0 * 64 + cc * 16, vi);
      ptx::tmem_ld16(t_lane + 1 * 64 + cc * 16, vj);
      ptx::tmem_ld16(t_lane + 2 * 64 + cc * 16, vf);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(crow + (((cc * 4 + q) ^ (r & 7)) << 4));
        cv[q * 4 + 0] = t.x; cv[q * 4 + 1] = t.y; cv[q * 4 + 2] = t.z; cv[q * 4 + 3] = t.w;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int j = cc * 16 + e;
        const float in = fmaf((vi[e] - mu[0]) * rs[0], ln.gamma[0][j], ln.beta[0][j]);
        const float jn = fmaf((vj[e] - mu[1]) * rs[1], ln.gamma[1][j], ln.beta[1][j]);
        const float fn = fmaf((vf[e] - mu[2]) * rs[2], ln.gamma[2][j], ln.beta[2][j]) + FORGET_BIAS;
        cn[j] = fmaf(cv[e], fast_sigmoid(fn), fast_sigmoid(in) * fmaxf(jn, 0.f));
      }
    }
    float cm, crs;
    {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) s += cn[j];
      cm = s * (1.0f / 64);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const float t = cn[j] - cm;
        q = fmaf(t, t, q);
      }
      crs = rsqrtf(q * (1.0f / 64) + LN_EPS);
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      float vo[16], hn[16];
      ptx::tmem_ld16(t_lane + 3 * 64 + cc * 16, vo);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int j = cc * 16 + e;
        const float on = fmaf((vo[e] - mu[3]) * rs[3], ln.gamma[3][j], ln.beta[3][j]);
        const float c2 = fmaf((cn[j] - cm) * crs, ln.gamma[4][j], ln.beta[4][j]);
        cn[j] = c2;
        hn[e] = fmaxf(c2, 0.f) * fast_sigmoid(on);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = cc * 16 + q * 4;
        *reinterpret_cast<float4*>(crow + (((cc * 4 + q) ^ (r & 7)) << 4)) =
            make_float4(cn[j], cn[j + 1], cn[j + 2], cn[j + 3]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {   // two 16-B chunks of 8 bf16
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) ptx::split_bf16x2(hn[q * 8 + 2 * p], hn[q * 8 + 2 * p + 1], hi[p], lo[p]);
        const uint32_t off = r * 128 + (((cc * 2 + q) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(hbuf + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (HP == 2) *reinterpret_cast<uint4*>(hbuf + IMG16_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    // ---- write the tile image back in place --------------------------------------------
    ptx::fence_proxy_async_smem();
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::bulk_s2g(gtile, hbuf, tile_bytes(HP));
      ptx::bulk_commit();
      ptx::bulk_wait_read0();
    }
    __syncthreads();
    phase ^= 1;
  }
  if (tid == 0) ptx::bulk_wait0();   // all tile images have landed in global memory
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

// ====================================================================================
// K2: message MLP chain (+ scatter / message store / vote)
// ====================================================================================
template <int HP>
struct K2Smem {
  static constexpr int W_BYTES = 4 * HP * 8192;            // [layer][plane] 8 KB images
  static constexpr int A_OFF = W_BYTES;
  static constexpr int B_OFF = A_OFF + HP * IMG16_BYTES;
  static constexpr int S_OFF = B_OFF + HP * IMG16_BYTES;   // fp32 staging for the scatter
  static constexpr int BAR_OFF = S_OFF + STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 64;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

template <int HP>
__global__ void __launch_bounds__(128, 1) tc_mlp_kernel(const K2Args a) {
  using L = K2Smem<HP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* bufs[2] = {smem + L::A_OFF, smem + L::B_OFF};
  float* stage = reinterpret_cast<float*>(smem + L::S_OFF);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* bar_ld = bar_w + 1;
  uint64_t* bar_mma = bar_w + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_v = static_cast<int>(blockIdx.x) >= a.e_ctas;
  const int bias_set = a.vote_mode ? 2 : (is_v ? 0 : 1);
  int t0, t1;
  if (is_v) tile_range(blockIdx.x - a.e_ctas, gridDim.x - a.e_ctas, a.tilesV, t0, t1);
  else tile_range(blockIdx.x, a.e_ctas, a.tilesE, t0, t1);
  const uint8_t* state = is_v ? a.stateV : a.stateE;
  const int64_t n_rows = is_v ? a.nV : a.nE;
  const int n_layers = a.vote_mode ? 3 : 4;

  if (tid == 0) {
    ptx::mbar_init(bar_w, 1);
    ptx::mbar_init(bar_ld, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 64);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0 && t0 < t1) {
    const uint8_t* wimg = is_v ? a.wV : a.wE;
    ptx::mbar_arrive_expect_tx(bar_w, L::W_BYTES);
    for (int off = 0; off < L::W_BYTES; off += 16384) ptx::bulk_g2s(wsm + off, wimg + off, 16384, bar_w);
  }

  constexpr uint32_t IDESC = ptx::umma_idesc_bf16(128, 64);
  const MlpBias& bias = c_mlp_bias[bias_set];
  const int r = tid;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t ld_phase = 0, mma_phase = 0;

  for (int tile = t0; tile < t1; ++tile) {
    const uint8_t* gtile = state + static_cast<int64_t>(tile) * tile_bytes(HP);
    const int64_t row0 = static_cast<int64_t>(tile) * TILE_ROWS;
    if (tid == 0) {
      ptx::mbar_arrive_expect_tx(bar_ld, HP * IMG16_BYTES);
      ptx::bulk_g2s(bufs[0], gtile, HP * IMG16_BYTES, bar_ld);
    }
    int cur = 0;
    float v[64];
    for (int l = 0; l < n_layers; ++l) {
      if (tid == 0) {
        if (l == 0) {
          if (tile == t0) ptx::mbar_wait(bar_w, 0);
          ptx::mbar_wait(bar_ld, ld_phase);
        }
        ptx::tcgen05_fence_after();
        uint32_t acc = 0;
        constexpr int NCOMB = (HP == 2) ? 3 : 1;
        const int pa_[3] = {1, 0, 0}, pb_[3] = {0, 1, 0};
#pragma unroll
        for (int cb = 0; cb < NCOMB; ++cb) {
          const int pa = (HP == 2) ? pa_[cb] : 0, pb = (HP == 2) ? pb_[cb] : 0;
          const uint32_t abase = ptx::smem_u32(bufs[cur] + pa * IMG16_BYTES);
          const uint32_t bbase = ptx::smem_u32(wsm + (l * HP + pb) * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ptx::umma_bf16_ss(tmem, ptx::umma_desc_k_sw128(abase + k * 32), ptx::umma_desc_k_sw128(bbase + k * 32),
                              IDESC, acc);
            acc = 1;
          }
        }
        ptx::umma_commit(bar_mma);
      }
      __syncwarp();
      ptx::mbar_wait(bar_mma, mma_phase);
      mma_phase ^= 1;
      ptx::tcgen05_fence_after();
      ptx::tmem_ld64(t_lane, v);
      const bool hidden = a.vote_mode || (l < 3);
      if (hidden) {
        uint8_t* nxt = bufs[cur ^ 1];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int j = ch * 8 + 2 * p;
            const float x0 = fmaxf(v[j] + bias.b[l][j], 0.f);
            const float x1 = fmaxf(v[j + 1] + bias.b[l][j + 1], 0.f);
            v[j] = x0;
            v[j + 1] = x1;
            ptx::split_bf16x2(x0, x1, hi[p], lo[p]);
          }
          const uint32_t off = r * 128 + ((ch ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(nxt + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (HP == 2) *reinterpret_cast<uint4*>(nxt + IMG16_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        ptx::fence_proxy_async_smem();
        cur ^= 1;
      }
      ptx::tcgen05_fence_before();
      __syncthreads();
    }
    ld_phase ^= 1;
    const int64_t grow = row0 + r;
    if (a.vote_mode) {
      // 64 -> 1 tail of E_vote on the fp32 layer-3 activations (model.py:107-128)
      float s = c_vote_tail.b4;
#pragma unroll
      for (int j = 0; j < 64; ++j) s = fmaf(v[j], c_vote_tail.w4[j], s);
      if (grow < n_rows) a.vote[grow] = s;
    } else if (is_v) {
      if (grow < n_rows) {
        float4* out = reinterpret_cast<float4*>(a.mV + grow * D);
#pragma unroll
        for (int q = 0; q < 16; ++q)
          out[q] = make_float4(v[4 * q] + bias.b[3][4 * q], v[4 * q + 1] + bias.b[3][4 * q + 1],
                               v[4 * q + 2] + bias.b[3][4 * q + 2], v[4 * q + 3] + bias.b[3][4 * q + 3]);
      }
    } else {
      // stage messages, then each warp walks its 32 rows: dst side one vector reduction per
      // row, src side accumulated over runs of equal src (rows are sorted by src)
      float* srow = stage + r * STAGE_LD;
#pragma unroll
      for (int q = 0; q < 32; ++q)
        *reinterpret_cast<float2*>(srow + 2 * q) =
            make_float2(v[2 * q] + bias.b[3][2 * q], v[2 * q + 1] + bias.b[3][2 * q + 1]);
      __syncwarp();   // rows of a warp are staged and consumed by the same warp
      const int64_t g0 = row0 + warp * 32;
      int my_s = -1, my_d = -1;
      if (g0 + lane < n_rows) {
        my_s = a.src[g0 + lane];
        my_d = a.dst[g0 + lane];
      }
      int cur_s = -1;
      float2 acc = make_float2(0.f, 0.f);
      for (int rr = 0; rr < 32; ++rr) {
        const int s = __shfl_sync(0xffffffffu, my_s, rr);
        const int d = __shfl_sync(0xffffffffu, my_d, rr);
        if (s < 0) break;
        const float2 m = *reinterpret_cast<const float2*>(stage + (warp * 32 + rr) * STAGE_LD + 2 * lane);
        ptx::red_add_v2(a.xV + static_cast<int64_t>(d) * D + 2 * lane, m.x, m.y);
        if (s != cur_s) {
          if (cur_s >= 0) ptx::red_add_v2(a.xV + static_cast<int64_t>(cur_s) * D + 2 * lane, acc.x, acc.y);
          cur_s = s;
          acc = m;
        } else {
          acc.x += m.x;
          acc.y += m.y;
        }
      }
      if (cur_s >= 0) ptx::red_add_v2(a.xV + static_cast<int64_t>(cur_s) * D + 2 * lane, acc.x, acc.y);
      __syncwarp();
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
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
    *reinterpret_cast<uint32_t*>(tile + img16_off(r, col)) = hi;
    if (HP == 2) *reinterpret_cast<uint32_t*>(tile + IMG16_BYTES + img16_off(r, col)) = lo;
  }
  if (c) {
    float2 cv = make_float2(0.f, 0.f);
    if (row < n_rows) cv = *reinterpret_cast<const float2*>(c + row * D + col);
    *reinterpret_cast<float2*>(tile + HP * IMG16_BYTES + img32_off(r, col)) = cv;
  }
}

// zero the c tile of every state tile (graphnn.py:137)
template <int HP>
__global__ void __launch_bounds__(256) tc_zero_c_kernel(int64_t n_rows_pad, uint8_t* __restrict__ state) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // float4 index
  if (i >= n_rows_pad * 16) return;
  const int64_t row = i >> 4;
  uint8_t* tile = state + (row / TILE_ROWS) * tile_bytes(HP) + HP * IMG16_BYTES;
  reinterpret_cast<float4*>(tile)[(row % TILE_ROWS) * 16 + (i & 15)] = make_float4(0.f, 0.f, 0.f, 0.f);
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
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(tile + img16_off(r, col));
    float x0 = __uint_as_float(hi << 16), x1 = __uint_as_float(hi & 0xFFFF0000u);
    if (HP == 2) {
      const uint32_t lo = *reinterpret_cast<const uint32_t*>(tile + IMG16_BYTES + img16_off(r, col));
      x0 += __uint_as_float(lo << 16);
      x1 += __uint_as_float(lo & 0xFFFF0000u);
    }
    *reinterpret_cast<float2*>(h + row * D + col) = make_float2(x0, x1);
  }
  if (c) *reinterpret_cast<float2*>(c + row * D + col) =
      *reinterpret_cast<const float2*>(tile + HP * IMG16_BYTES + img32_off(r, col));
}

}  // namespace tspgnn
