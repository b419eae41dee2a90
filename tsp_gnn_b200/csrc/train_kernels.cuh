// fp32 CUDA-core kernels of the training step (model.py:157-167): the reverse pass through the
// message-passing loop, the vote / initial-embedding MLPs, and the optimizer.
//
// The reverse pass recomputes each timestep's intermediates from the (h, c) snapshots the training
// forward keeps, so it is built from a few generic pieces on row-major fp32 [rows, 64*k] matrices:
//   rowgemm_kernel   Y = epi(X . W (+ b))           X in 64-column blocks, W or W^T resident in smem
//   xtdy_kernel      dW += X^T . dY, db += colsum(dY)   (reduction over rows, fp32 atomics)
//   lstm_bwd_kernel  reverse of the LayerNorm-LSTM gate math, one warp per row
//   gather2 / scatter2   EV . y  and  EV^T . y  for the two non-zeros of every edge row
// All matrices are addressed as (pointer, leading dimension); rows >= n_rows are never touched.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace tspgnn {

constexpr int RG_THREADS = 256;            // threads per CTA = rows per tile of rowgemm
constexpr int RG_XLD = 65;                 // row-major staging [row][64], +1 breaks bank conflicts
constexpr int RG_MAXB = 4;                 // up to 4 blocks of 64 columns on either side

struct RowGemmArgs {
  const float* x[RG_MAXB];   // input column block kb: x[kb][row * xld[kb] + 0..63]
  int xld[RG_MAXB];
  float* y[RG_MAXB];         // output column block nb
  int yld[RG_MAXB];
  const float* w;            // stored matrix Ws[i * ldw + j], valid for i < w_rows, j < w_cols (else 0)
  int ldw, w_rows, w_cols;
  const float* bias;         // [N] (first bias_n valid) or nullptr
  int bias_n;
  const float* mask;         // EPI_MASK: multiply by (mask[row * mld + col] > 0); one 64-column block (N == 64)
  int mld;
  int64_t n_rows;
  long long* timeline;       // optional clock64() trace of tc_rowgemm_kernel's roles (tools/timeline_train.py), nullptr in production
};

constexpr int EPI_NONE = 0, EPI_RELU = 1, EPI_MASK = 2, EPI_ACCUM = 4;

__host__ __device__ constexpr int rowgemm_smem_bytes(int K, int N, int RPT) { return (K * N + 32 * RPT * RG_XLD) * 4; }

// Y[r, n] = epi( sum_k X[r, k] * Wm[k, n] + bias[n] ),  K = 64*KB, N = 64*NB.
// TRANS = false: Wm[k, n] = Ws[k, n];  TRANS = true: Wm[k, n] = Ws[n, k]  (dX = dY . W^T).
// EPI flags: RELU, MASK (zero where the forward activation was not positive), ACCUM (Y += ...).
// A CTA owns tiles of 32*RPT rows (RPT = 8: 256 rows for the edge-sized matrices; RPT = 2: 64 rows, so
// that the vertex-sized ones still spread over 80 CTAs instead of 20); per (tile, 64-column block)
// thread (rg = tid / 8, cg = tid % 8) accumulates the RPT x 8 micro-tile rows rg*RPT .. x columns
// {cg*4 .. +3, 32 + cg*4 .. +3}.  With RPT = 8, per k-step eight
// LDS.32 of x (four distinct rows per warp, broadcast) and two conflict-free LDS.128 of W feed 64 FMAs,
// which balances the shared-memory pipe against the FMA pipe (a one-row-per-thread form needs 16
// LDS.128 per 64 FMAs and is bound by shared-memory bandwidth at a quarter of the FMA rate).
template <int KB, int NB, bool TRANS, int EPI, int RPT>
__global__ void __launch_bounds__(RG_THREADS) rowgemm_kernel(const RowGemmArgs a) {
  extern __shared__ float smem[];
  constexpr int K = 64 * KB, N = 64 * NB;
  constexpr int TILE_R = 32 * RPT;                        // rows per tile
  constexpr int PASSES = (RPT == 8) ? 2 : 1;              // staging passes
  constexpr int PER = TILE_R * 16 / RG_THREADS / PASSES;  // float4 per thread and pass
  float* Wm = smem;              // [K][N]
  float* xs = smem + K * N;      // [TILE_R][RG_XLD]: the X operand, then the output block
  const int tid = threadIdx.x;
  const int rg = tid >> 3, cg = tid & 7;
  // Weights: 16-byte loads along the rows of the stored matrix (every ldw / w_cols in use is a
  // multiple of 4 and every matrix starts 16-byte aligned in the blob).
  if (!TRANS) {
#pragma unroll 4
    for (int i4 = tid; i4 < K * N / 4; i4 += RG_THREADS) {
      const int k = i4 / (N / 4), n = (i4 % (N / 4)) * 4;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < a.w_rows && n < a.w_cols) w = *reinterpret_cast<const float4*>(a.w + static_cast<int64_t>(k) * a.ldw + n);
      *reinterpret_cast<float4*>(Wm + k * N + n) = w;
    }
  } else {
    // consecutive threads take consecutive n: conflict-free transposed stores
#pragma unroll 4
    for (int i4 = tid; i4 < K * N / 4; i4 += RG_THREADS) {
      const int n = i4 % N, k = (i4 / N) * 4;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < a.w_rows && k < a.w_cols) w = *reinterpret_cast<const float4*>(a.w + static_cast<int64_t>(n) * a.ldw + k);
      Wm[(k + 0) * N + n] = w.x;
      Wm[(k + 1) * N + n] = w.y;
      Wm[(k + 2) * N + n] = w.z;
      Wm[(k + 3) * N + n] = w.w;
    }
  }
  const int64_t n_tiles = (a.n_rows + TILE_R - 1) / TILE_R;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * TILE_R;
#pragma unroll 1
    for (int nb = 0; nb < NB; ++nb) {
      float acc[RPT][8];
      {
        float b8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = nb * 64 + (j >> 2) * 32 + cg * 4 + (j & 3);
          b8[j] = (a.bias != nullptr && col < a.bias_n) ? a.bias[col] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = b8[j];
      }
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
        __syncthreads();   // previous users of xs (and the weight load) are done
        const float* xp = a.x[kb];
        const int xld = a.xld[kb];
        // TILE_R rows x 16 float4, coalesced; PER loads in flight per thread and pass
#pragma unroll 1
        for (int pass = 0; pass < PASSES; ++pass) {
          float4 xv[PER];
#pragma unroll
          for (int q = 0; q < PER; ++q) {
            const int i4 = tid + (pass * PER + q) * RG_THREADS;
            const int r = i4 >> 4, c4 = i4 & 15;
            xv[q] = (row0 + r < a.n_rows) ? *reinterpret_cast<const float4*>(xp + (row0 + r) * xld + c4 * 4)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int q = 0; q < PER; ++q) {
            const int i4 = tid + (pass * PER + q) * RG_THREADS;
            const int r = i4 >> 4, c4 = i4 & 15;
            float* d = xs + r * RG_XLD + c4 * 4;
            d[0] = xv[q].x; d[1] = xv[q].y; d[2] = xv[q].z; d[3] = xv[q].w;
          }
        }
        __syncthreads();
        const float* Wk = Wm + (kb * 64) * N + nb * 64 + cg * 4;
        const float* xr = xs + (rg * RPT) * RG_XLD;
#pragma unroll 4
        for (int k = 0; k < 64; ++k) {
          const float4 w0 = *reinterpret_cast<const float4*>(Wk + k * N);
          const float4 w1 = *reinterpret_cast<const float4*>(Wk + k * N + 32);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const float xv = xr[i * RG_XLD + k];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
          }
        }
      }
      __syncthreads();   // everyone is done reading xs as the X operand
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        float* d = xs + (rg * RPT + i) * RG_XLD + cg * 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = (EPI & EPI_RELU) ? fmaxf(acc[i][j], 0.f) : acc[i][j];
          d[(j >> 2) * 32 + (j & 3)] = v;
        }
      }
      __syncthreads();
      float* yp = a.y[nb];
      const int yld = a.yld[nb];
#pragma unroll 1
      for (int pass = 0; pass < PASSES; ++pass) {
        float4 mv[PER], ov[PER];
        if (EPI & (EPI_MASK | EPI_ACCUM)) {
#pragma unroll
          for (int q = 0; q < PER; ++q) {
            const int i4 = tid + (pass * PER + q) * RG_THREADS;
            const int r = i4 >> 4, c4 = i4 & 15;
            mv[q] = ov[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < a.n_rows) {
              if (EPI & EPI_MASK) mv[q] = *reinterpret_cast<const float4*>(a.mask + (row0 + r) * a.mld + c4 * 4);
              if (EPI & EPI_ACCUM) ov[q] = *reinterpret_cast<const float4*>(yp + (row0 + r) * yld + c4 * 4);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int i4 = tid + (pass * PER + q) * RG_THREADS;
          const int r = i4 >> 4, c4 = i4 & 15;
          if (row0 + r < a.n_rows) {
            const float* sp = xs + r * RG_XLD + c4 * 4;
            float4 v = make_float4(sp[0], sp[1], sp[2], sp[3]);
            if (EPI & EPI_MASK) {
              v.x = (mv[q].x > 0.f) ? v.x : 0.f;
              v.y = (mv[q].y > 0.f) ? v.y : 0.f;
              v.z = (mv[q].z > 0.f) ? v.z : 0.f;
              v.w = (mv[q].w > 0.f) ? v.w : 0.f;
            }
            if (EPI & EPI_ACCUM) {
              v.x += ov[q].x; v.y += ov[q].y; v.z += ov[q].z; v.w += ov[q].w;
            }
            *reinterpret_cast<float4*>(yp + (row0 + r) * yld + c4 * 4) = v;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// dW[kb*64 + k, nb*64 + n] += sum_r X[r, kb*64 + k] * dY[r, nb*64 + n]      (blockIdx.y = kb, blockIdx.z = nb)
// db[nb*64 + n]            += sum_r dY[r, nb*64 + n]                          (only the kb == 0 CTAs)
// 256 threads = 4 row groups x 64 threads; a group takes every fourth row of the 64-row chunks
// streamed through shared memory, and each of its threads an 8 x 8 block of the 64 x 64 tile
// (k = ty*8 .. +7, n = {tx*4 .. +3, 32 + tx*4 .. +3}): four conflict-free LDS.128 feed 64 FMAs.
// The four partial tiles are summed in shared memory, then one fp32 atomic per element goes out.
// ------------------------------------------------------------------------------------
struct XtdyArgs {
  const float* x[RG_MAXB];
  int xld[RG_MAXB];
  const float* dy;      // dY[row * dyld + nb*64 + n]
  int dyld;
  float* dw;            // dW[i * ldw + j], only i < w_rows, j < w_cols are written
  int ldw, w_rows, w_cols;
  float* db;            // nullptr or [w_cols]
  int64_t n_rows;
};

constexpr int XT_ROWS = 64;
constexpr int XT_SMEM_BYTES = 2 * 2 * XT_ROWS * 64 * 4;   // two stages x (X chunk, dY chunk)

__device__ __forceinline__ void cp_async16_zfill(uint32_t saddr, const void* gptr, bool valid) {
  // 16-byte asynchronous copy global -> shared; src-size 0 zero-fills the destination (rows past the end)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gptr), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256) xtdy_kernel(const XtdyArgs a) {
  extern __shared__ __align__(16) float xt_smem[];   // [stage][X | dY][64 rows][64]
  const int tid = threadIdx.x;
  const int g = tid >> 6, t = tid & 63;
  const int ty = t >> 3, tx = t & 7;
  const int kb = blockIdx.y, nb = blockIdx.z;
  const float* xp = a.x[kb];
  const int xld = a.xld[kb];
  const float* yp = a.dy + nb * 64;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool do_bias = (a.db != nullptr) && kb == 0 && ty == 0;
  const int64_t n_chunks = (a.n_rows + XT_ROWS - 1) / XT_ROWS;
  const uint32_t smem_s = ptx::smem_u32(xt_smem);
  // stage `st` <- chunk `ch`: 64 rows x 16 float4 per matrix, four asynchronous 16-byte copies per thread and matrix
  auto prefetch = [&](int64_t ch, int st) {
    const int64_t row0 = ch * XT_ROWS;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = tid + q * 256;
      const int r = i >> 4, c4 = i & 15;
      const bool valid = row0 + r < a.n_rows;
      const int64_t row = valid ? row0 + r : 0;
      const uint32_t dst = smem_s + ((st * 2) * XT_ROWS * 64 + r * 64 + c4 * 4) * 4;
      cp_async16_zfill(dst, xp + row * xld + c4 * 4, valid);
      cp_async16_zfill(dst + XT_ROWS * 64 * 4, yp + row * a.dyld + c4 * 4, valid);
    }
    cp_async_commit();
  };
  int64_t ch = blockIdx.x;
  if (ch < n_chunks) prefetch(ch, 0);
  for (int it = 0; ch < n_chunks; ch += gridDim.x, ++it) {
    const int st = it & 1;
    const bool more = ch + gridDim.x < n_chunks;
    if (more) prefetch(ch + gridDim.x, st ^ 1);      // overlaps the math of this chunk
    if (more) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const float* Xs = xt_smem + (st * 2) * XT_ROWS * 64;
    const float* Ys = Xs + XT_ROWS * 64;
#pragma unroll 4
    for (int rr = 0; rr < XT_ROWS / 4; ++rr) {
      const int r = rr * 4 + g;
      const float4 x0 = *reinterpret_cast<const float4*>(Xs + r * 64 + ty * 8);
      const float4 x1 = *reinterpret_cast<const float4*>(Xs + r * 64 + ty * 8 + 4);
      const float4 y0 = *reinterpret_cast<const float4*>(Ys + r * 64 + tx * 4);
      const float4 y1 = *reinterpret_cast<const float4*>(Ys + r * 64 + 32 + tx * 4);
      const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const float ya[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bs[j] += ya[j];
      }
    }
    __syncthreads();   // this stage is overwritten by the prefetch issued in the next iteration
  }
  // sum the four row groups in shared memory (re-using the staging buffers), then one atomic per element
  float* red = xt_smem;                 // [64][64]
  float* redb = xt_smem + 64 * 64;      // [64]
  for (int i = tid; i < 64 * 64 + 64; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      atomicAdd(red + (ty * 8 + i) * 64 + (j >> 2) * 32 + tx * 4 + (j & 3), acc[i][j]);
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(redb + (j >> 2) * 32 + tx * 4 + (j & 3), bs[j]);
  }
  __syncthreads();
  for (int i = tid; i < 64 * 64; i += 256) {
    const int wi = kb * 64 + (i >> 6), wj = nb * 64 + (i & 63);
    if (wi < a.w_rows && wj < a.w_cols) atomicAdd(a.dw + static_cast<int64_t>(wi) * a.ldw + wj, red[i]);
  }
  if (a.db != nullptr && kb == 0 && tid < 64) {
    const int wj = nb * 64 + tid;
    if (wj < a.w_cols) atomicAdd(a.db + wj, redb[tid]);
  }
}

// ------------------------------------------------------------------------------------
// Reverse of the LayerNorm-LSTM gate math, one warp per row, lane = columns
// lane and lane + 32 of every gate.
//   in : z [rows,256] pre-LayerNorm gates (recomputed), c_prev, g_h = dL/dh', g_c = dL/dc'
//   out: z  <- dL/dz (in place),  g_c <- dL/dc_prev (in place),
//        grad blob: += dL/dgamma, dL/dbeta of the five LayerNorms (fp32 atomics)
// ------------------------------------------------------------------------------------
struct LnOffsets {
  int gamma[5], beta[5];   // offsets into the parameter / gradient blob; order input, transform, forget, output, state
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// the LSTM reverse kernel's form: ex2 + rcp (about 1e-6 relative, far inside the gradient gates), a third of the instructions
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// forward LayerNorm statistics of a 64-vector held as 2 values per lane: uh = (u - mean) * r
__device__ __forceinline__ void ln_fwd2(const float (&u)[2], float (&uh)[2], float& r) {
  const float mean = warp_sum(u[0] + u[1]) * (1.0f / 64.0f);
  const float d0 = u[0] - mean, d1 = u[1] - mean;
  const float var = warp_sum(d0 * d0 + d1 * d1) * (1.0f / 64.0f);
  r = rsqrtf(var + LN_EPS);
  uh[0] = d0 * r;
  uh[1] = d1 * r;
}

// reverse of y = uh * gamma + beta: returns du, accumulates dgamma / dbeta
__device__ __forceinline__ void ln_bwd2(const float (&dy)[2], const float (&uh)[2], float r, const float (&gamma)[2],
                                        float (&du)[2], float (&dgam)[2], float (&dbet)[2]) {
  dgam[0] += dy[0] * uh[0];
  dgam[1] += dy[1] * uh[1];
  dbet[0] += dy[0];
  dbet[1] += dy[1];
  const float a0 = dy[0] * gamma[0], a1 = dy[1] * gamma[1];
  const float m1 = warp_sum(a0 + a1) * (1.0f / 64.0f);
  const float m2 = warp_sum(a0 * uh[0] + a1 * uh[1]) * (1.0f / 64.0f);
  du[0] = r * (a0 - m1 - uh[0] * m2);
  du[1] = r * (a1 - m1 - uh[1] * m2);
}

// Half-warp sums (lanes 0-15 and 16-31 each reduce their own values), NS independent scalars at once so that the
// shuffle chains of several reductions overlap.
template <int NS>
__device__ __forceinline__ void hw_sum(float (&v)[NS]) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < NS; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
}

// reverse of y = uh * gamma + beta for a 64-vector held as 4 values per lane of a half-warp
__device__ __forceinline__ void ln_bwd4(const float (&dy)[4], const float (&uh)[4], float r, const float (&gamma)[4],
                                        float (&du)[4], float (&dgam)[4], float (&dbet)[4]) {
  float a[4], m[2] = {0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    dgam[q] += dy[q] * uh[q];
    dbet[q] += dy[q];
    a[q] = dy[q] * gamma[q];
    m[0] += a[q];
    m[1] = fmaf(a[q], uh[q], m[1]);
  }
  hw_sum<2>(m);
  const float m1 = m[0] * (1.0f / 64.0f), m2 = m[1] * (1.0f / 64.0f);
#pragma unroll
  for (int q = 0; q < 4; ++q) du[q] = r * (a[q] - m1 - uh[q] * m2);
}

// One HALF-WARP per row: lane hl = lane & 15 holds columns 4 hl .. 4 hl + 3 of every 64-vector of the row, so every
// access is a 16-byte load / store (a row of z is four of them per lane), a reduction takes four shuffle steps that
// serve two rows at once (40 shuffles per row instead of 100 with a warp per row and 4-byte accesses), and the four
// gate LayerNorms reduce together.  z is overwritten with dz, g_c with dL/dc.
__global__ void __launch_bounds__(256, 2) lstm_bwd_kernel(float* __restrict__ z, const float* __restrict__ c_prev,
                                                       const float* __restrict__ g_h, float* __restrict__ g_c,
                                                       int64_t n_rows, const float* __restrict__ params,
                                                       float* __restrict__ grads, const LnOffsets off) {
  const int lane = threadIdx.x & 31, hl = lane & 15;
  const int64_t hw = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 4;
  const int64_t n_hw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 4;
  // LayerNorm parameters: shared memory (read per row as 16-byte pieces), the 40 gradient sums stay in registers
  __shared__ __align__(16) float prm[2 * 5 * 64];
  __shared__ float red[2 * 5 * 64];
  for (int i = threadIdx.x; i < 2 * 5 * 64; i += blockDim.x) {
    const int g = (i % 320) >> 6, col = i & 63;
    prm[i] = params[(i < 320 ? off.gamma[g] : off.beta[g]) + col];
    red[i] = 0.f;
  }
  __syncthreads();
  float dgam[5][4], dbet[5][4];
#pragma unroll
  for (int g = 0; g < 5; ++g)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dgam[g][q] = 0.f;
      dbet[g][q] = 0.f;
    }
  auto ld4 = [&](const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * hl);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  };
  // both halves of a warp run the same number of iterations (the shuffles are warp-wide): rows past the end are
  // computed on zeros and never stored
  const int64_t n_iter = (n_rows + n_hw - 1) / n_hw;
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t row = hw + it * n_hw;
    const bool live = row < n_rows;
    float zg[4][4], uh[4][4], r[4], c[4], gh[4], gc[4];
    {
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 t[7];
#pragma unroll
      for (int g = 0; g < 4; ++g) t[g] = live ? *reinterpret_cast<const float4*>(z + row * 256 + g * 64 + 4 * hl) : zero;
      t[4] = live ? *reinterpret_cast<const float4*>(c_prev + row * 64 + 4 * hl) : zero;
      t[5] = live ? *reinterpret_cast<const float4*>(g_h + row * 64 + 4 * hl) : zero;
      t[6] = live ? *reinterpret_cast<const float4*>(g_c + row * 64 + 4 * hl) : zero;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        zg[g][0] = t[g].x; zg[g][1] = t[g].y; zg[g][2] = t[g].z; zg[g][3] = t[g].w;
      }
      c[0] = t[4].x; c[1] = t[4].y; c[2] = t[4].z; c[3] = t[4].w;
      gh[0] = t[5].x; gh[1] = t[5].y; gh[2] = t[5].z; gh[3] = t[5].w;
      gc[0] = t[6].x; gc[1] = t[6].y; gc[2] = t[6].z; gc[3] = t[6].w;
    }
    // forward LayerNorm statistics of the four gates, reduced together
    {
      float m[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) m[g] = (zg[g][0] + zg[g][1]) + (zg[g][2] + zg[g][3]);
      hw_sum<4>(m);
      float v[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float mean = m[g] * (1.0f / 64.0f);
        v[g] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uh[g][q] = zg[g][q] - mean;
          v[g] = fmaf(uh[g][q], uh[g][q], v[g]);
        }
      }
      hw_sum<4>(v);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        r[g] = rsqrtf(v[g] * (1.0f / 64.0f) + LN_EPS);
#pragma unroll
        for (int q = 0; q < 4; ++q) uh[g][q] *= r[g];
      }
    }
    float si[4], sf[4], so[4], jj[4], gg[4], ct[4];
    {
      float gm[4][4], bt[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        ld4(prm + g * 64, gm[g]);
        ld4(prm + 320 + g * 64, bt[g]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        si[q] = sigmoidf_fast(uh[0][q] * gm[0][q] + bt[0][q]);
        jj[q] = uh[1][q] * gm[1][q] + bt[1][q];
        gg[q] = fmaxf(jj[q], 0.f);
        sf[q] = sigmoidf_fast(uh[2][q] * gm[2][q] + bt[2][q] + FORGET_BIAS);
        so[q] = sigmoidf_fast(uh[3][q] * gm[3][q] + bt[3][q]);
        ct[q] = c[q] * sf[q] + si[q] * gg[q];
      }
    }
    float ch[4], cr;
    {
      float m[1] = {(ct[0] + ct[1]) + (ct[2] + ct[3])};
      hw_sum<1>(m);
      const float mean = m[0] * (1.0f / 64.0f);
      float v[1] = {0.f};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ch[q] = ct[q] - mean;
        v[0] = fmaf(ch[q], ch[q], v[0]);
      }
      hw_sum<1>(v);
      cr = rsqrtf(v[0] * (1.0f / 64.0f) + LN_EPS);
#pragma unroll
      for (int q = 0; q < 4; ++q) ch[q] *= cr;
    }
    float d_so[4], d_cn[4], d_ct[4], gmv[4];
    {
      float bt[4];
      ld4(prm + 4 * 64, gmv);
      ld4(prm + 320 + 4 * 64, bt);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float cn = ch[q] * gmv[q] + bt[q];
        d_so[q] = gh[q] * fmaxf(cn, 0.f);
        d_cn[q] = gc[q] + ((cn > 0.f) ? gh[q] * so[q] : 0.f);
      }
    }
    ln_bwd4(d_cn, ch, cr, gmv, d_ct, dgam[4], dbet[4]);
    float dgate[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dgate[0][q] = d_ct[q] * gg[q] * si[q] * (1.0f - si[q]);
      dgate[1][q] = (jj[q] > 0.f) ? d_ct[q] * si[q] : 0.f;
      dgate[2][q] = d_ct[q] * c[q] * sf[q] * (1.0f - sf[q]);
      dgate[3][q] = d_so[q] * so[q] * (1.0f - so[q]);
    }
    if (live)
      *reinterpret_cast<float4*>(g_c + row * 64 + 4 * hl) =
          make_float4(d_ct[0] * sf[0], d_ct[1] * sf[1], d_ct[2] * sf[2], d_ct[3] * sf[3]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float du[4];
      ld4(prm + g * 64, gmv);
      ln_bwd4(dgate[g], uh[g], r[g], gmv, du, dgam[g], dbet[g]);
      if (live) *reinterpret_cast<float4*>(z + row * 256 + g * 64 + 4 * hl) = make_float4(du[0], du[1], du[2], du[3]);
    }
  }
  // LayerNorm parameter gradients: summed over the CTA's half-warps in shared memory, then one global
  // atomic per parameter and CTA (640 addresses would otherwise serialise every warp of the grid)
#pragma unroll
  for (int g = 0; g < 5; ++g)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      atomicAdd(red + g * 64 + 4 * hl + q, dgam[g][q]);
      atomicAdd(red + 320 + g * 64 + 4 * hl + q, dbet[g][q]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 5 * 64; i += blockDim.x) {
    const int g = (i % 320) >> 6, col = i & 63;
    atomicAdd(grads + (i < 320 ? off.gamma[g] : off.beta[g]) + col, red[i]);
  }
}

// ------------------------------------------------------------------------------------
// out[e, :] = in[src[e], :] + in[dst[e], :]    ( = EV . in, graphnn.py:156-160 )
// 16 threads per row, one float4 each.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather2_kernel(const float* __restrict__ in, const int32_t* __restrict__ src,
                                                      const int32_t* __restrict__ dst, int64_t n_rows,
                                                      float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows * 16) return;
  const int64_t e = i >> 4;
  const int c4 = static_cast<int>(i & 15);
  const float4 u = reinterpret_cast<const float4*>(in + static_cast<int64_t>(src[e]) * 64)[c4];
  const float4 w = reinterpret_cast<const float4*>(in + static_cast<int64_t>(dst[e]) * 64)[c4];
  reinterpret_cast<float4*>(out + e * 64)[c4] = make_float4(u.x + w.x, u.y + w.y, u.z + w.z, u.w + w.w);
}

// out[src[e], :] += in[e, :]; out[dst[e], :] += in[e, :]    ( = EV^T . in ); `out` must be zeroed first.
// `in` has leading dimension ld.  16 threads per row, one red.global.add.v4.f32 per endpoint.
__global__ void __launch_bounds__(256) scatter2_kernel(const float* __restrict__ in, int ld,
                                                       const int32_t* __restrict__ src,
                                                       const int32_t* __restrict__ dst, int64_t n_rows,
                                                       float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows * 16) return;
  const int64_t e = i >> 4;
  const int c4 = static_cast<int>(i & 15);
  const float4 v = *reinterpret_cast<const float4*>(in + e * ld + c4 * 4);
  ptx::red_add_v4(out + static_cast<int64_t>(src[e]) * 64 + c4 * 4, v);
  ptx::red_add_v4(out + static_cast<int64_t>(dst[e]) * 64 + c4 * 4, v);
}

// ------------------------------------------------------------------------------------
// Loss (model.py:157) and its gradient with respect to the per-edge votes:
//   loss = (1/Bg) sum_k max(l,0) - l*y + log1p(exp(-|l|));  dvote_k = (sigmoid(l_k) - y_k) / (Bg * n_edges[k])
// Bg = global batch (the divisor of reduce_mean when instances are sharded over ranks).  One CTA.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ logits,
                                                        const float* __restrict__ route_exists,
                                                        const int64_t* __restrict__ eoff, int n_instances,
                                                        float inv_global_batch, float* __restrict__ dvote_inst,
                                                        float* __restrict__ loss) {
  __shared__ float part[8];
  float s = 0.f;
  for (int k = threadIdx.x; k < n_instances; k += blockDim.x) {
    const float l = logits[k], y = route_exists[k];
    s += fmaxf(l, 0.f) - l * y + log1pf(expf(-fabsf(l)));
    const float ne = static_cast<float>(eoff[k + 1] - eoff[k]);
    dvote_inst[k] = (sigmoidf_acc(l) - y) * inv_global_batch / ne;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w];
    *loss = t * inv_global_batch;
  }
}

// Reverse of the 64->1 tail of E_vote (model.py:107-128): vote[e] = a3[e,:] . w4 + b4.
//   d3[e, j] = dvote[inst(e)] * w4[j] * (a3[e, j] > 0);  dw4[j] += sum_e a3[e, j] dvote;  db4 += sum_e dvote
// One warp per row (grid-stride), lane = columns lane, lane + 32.
__global__ void __launch_bounds__(256) vote_tail_bwd_kernel(const float* __restrict__ a3,
                                                            const float* __restrict__ dvote_inst,
                                                            const int64_t* __restrict__ eoff, int n_instances,
                                                            int64_t n_rows, const float* __restrict__ w4,
                                                            float* __restrict__ d3, float* __restrict__ dw4,
                                                            float* __restrict__ db4) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const float w0 = w4[lane], w1 = w4[lane + 32];
  float g0 = 0.f, g1 = 0.f, gb = 0.f;
  for (int64_t row = warp; row < n_rows; row += n_warps) {
    // instance of this row: largest k with eoff[k] <= row
    int lo = 0, hi = n_instances - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (eoff[mid] <= row) lo = mid; else hi = mid - 1;
    }
    const float dv = dvote_inst[lo];
    const float x0 = a3[row * 64 + lane], x1 = a3[row * 64 + lane + 32];
    d3[row * 64 + lane] = (x0 > 0.f) ? dv * w0 : 0.f;
    d3[row * 64 + lane + 32] = (x1 > 0.f) ? dv * w1 : 0.f;
    g0 = fmaf(x0, dv, g0);
    g1 = fmaf(x1, dv, g1);
    gb += dv;
  }
  atomicAdd(dw4 + lane, g0);
  atomicAdd(dw4 + lane + 32, g1);
  if (lane == 0) atomicAdd(db4, gb);
}

// X0[e, :] = [W[e], C[e], 0, ...]  : the input of E_init_MLP (model.py:43) padded to 64 columns
__global__ void __launch_bounds__(256) pad_wc_kernel(const float* __restrict__ W, const float* __restrict__ C,
                                                     int64_t n_rows, float* __restrict__ X0) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows * 64) return;
  const int64_t e = i >> 6;
  const int c = static_cast<int>(i & 63);
  X0[i] = (c == 0) ? W[e] : (c == 1 ? C[e] : 0.f);
}

// dV_init[j] += sum_v gVh[v, j] / sqrt(d)     (V0 = tile(V_init / sqrt(d)), model.py:46-51)
__global__ void __launch_bounds__(256) vinit_grad_kernel(const float* __restrict__ gVh, int64_t n_rows, float scale,
                                                         float* __restrict__ dvinit) {
  const int j = threadIdx.x & 63;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 6);
  float s = 0.f;
  for (int64_t r = r0; r < n_rows; r += static_cast<int64_t>(gridDim.x) * 4) s += gVh[r * 64 + j];
  atomicAdd(dvinit + j, s * scale);
}

// ------------------------------------------------------------------------------------
// Optimizer (model.py:160-167): g += l2 * var;  global norm;  clip;  Adam.
// scal[0] = global norm of the (L2-augmented) gradient.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) grad_l2_norm_kernel(float* __restrict__ g, const float* __restrict__ p,
                                                            int64_t n, float l2, float* __restrict__ scal) {
  __shared__ float part[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = fmaf(l2, p[i], g[i]);
    g[i] = v;
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = part[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) scal[0] = sqrtf(t);
  }
}

// TF: clip_by_global_norm scales by clip / max(norm, clip);  AdamOptimizer._apply_dense:
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; var -= lr_t * m / (sqrt(v) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   const float* __restrict__ scal, float clip, float lr_t, float b1,
                                                   float b2, float eps) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float scale = clip / fmaxf(scal[0], clip);
  const float gi = g[i] * scale;
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

}  // namespace tspgnn
