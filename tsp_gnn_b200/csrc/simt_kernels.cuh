// fp32 CUDA-core kernels: the bring-up / cross-check path (TSPGNN_MODE_SIMT_FP32) and the
// once-per-forward pieces (edge-init MLP, vote tail, readout) shared by every mode.
// State layout here is plain row-major fp32 [rows, 64].
#pragma once
#include "common.cuh"

namespace tspgnn {

__constant__ CellLN c_ln[2];       // [0] = V cell, [1] = E cell
__constant__ MlpBias c_mlp_bias[3];  // [0] = V_msg_E, [1] = E_msg_V, [2] = E_vote (layers 1-3)
__constant__ VoteTail c_vote_tail;
__constant__ EInit c_einit;
__constant__ float c_vinit[D];     // V_init / sqrt(d)   (model.py:46-51)

constexpr int SIMT_THREADS = 128;
constexpr int XS_LD = SIMT_THREADS + 1;   // transposed staging [k][row], +1 breaks bank conflicts

// ---------------------------------------------------------------------------------
// E0 = E_init_MLP([W, C])   (model.py:33-43); one thread per edge row.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) simt_edge_init_kernel(const float* __restrict__ W,
                                                             const float* __restrict__ C, int64_t n_rows,
                                                             float* __restrict__ Eh) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  float in0 = W[r], in1 = C[r];
  float a1[8], a2[16], a3[32];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    a1[j] = fmaxf(fmaf(in1, c_einit.w1[1][j], fmaf(in0, c_einit.w1[0][j], c_einit.b1[j])), 0.f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float s = c_einit.b2[j];
#pragma unroll
    for (int k = 0; k < 8; ++k) s = fmaf(a1[k], c_einit.w2[k][j], s);
    a2[j] = fmaxf(s, 0.f);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float s = c_einit.b3[j];
#pragma unroll
    for (int k = 0; k < 16; ++k) s = fmaf(a2[k], c_einit.w3[k][j], s);
    a3[j] = fmaxf(s, 0.f);
  }
  float4* out = reinterpret_cast<float4*>(Eh + r * D);
#pragma unroll
  for (int j4 = 0; j4 < 16; ++j4) {
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int j = j4 * 4 + q;
      float s = c_einit.b4[j];
#pragma unroll
      for (int k = 0; k < 32; ++k) s = fmaf(a3[k], c_einit.w4[k][j], s);
      o[q] = s;
    }
    out[j4] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// V0 rows all equal V_init/sqrt(d) (model.py:46-51)
__global__ void simt_vertex_init_kernel(int64_t n_rows, float* __restrict__ Vh) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rows * D) Vh[i] = c_vinit[i % D];
}

// ---------------------------------------------------------------------------------
// y = L4(relu(L3(relu(L2(relu(L1(x)))))))  on rows x 64   (mlp.py:57-63, graphnn.py:114-125)
// weights: 4 x [64 in][64 out] fp32 row-major (tf.layers.Dense kernel layout).
// MODE 0: y is [rows,64]; MODE 1 (vote): layers 1-3 then the 64->1 tail, y is [rows].
// One thread per row, activations staged transposed in shared memory, weights in shared.
// ---------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(SIMT_THREADS) simt_mlp4_kernel(const float* __restrict__ x, int64_t n_rows,
                                                                 const float* __restrict__ weights, int layer_stride,
                                                                 int bias_set, float* __restrict__ y) {
  extern __shared__ float smem[];
  constexpr int NL = (MODE == 0) ? 4 : 3;
  float* Ws = smem;                       // [NL][64][64]
  float* xs = smem + NL * D * D;          // [64][XS_LD]
  const int tid = threadIdx.x;
  // layer l's kernel starts at weights + l*layer_stride (kernel, bias, kernel, ... in the blob)
  for (int i = tid; i < NL * D * D; i += SIMT_THREADS)
    Ws[i] = weights[(i / (D * D)) * layer_stride + (i % (D * D))];
  const int64_t n_tiles = (n_rows + SIMT_THREADS - 1) / SIMT_THREADS;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * SIMT_THREADS;
    __syncthreads();
    // coalesced load of the tile, transposed into xs[k][r]
    for (int i = tid; i < SIMT_THREADS * D; i += SIMT_THREADS) {
      int r = i / D, k = i % D;
      xs[k * XS_LD + r] = (row0 + r < n_rows) ? x[(row0 + r) * D + k] : 0.f;
    }
    __syncthreads();
    float acc[D];
#pragma unroll 1
    for (int l = 0; l < NL; ++l) {
#pragma unroll
      for (int j = 0; j < D; ++j) acc[j] = c_mlp_bias[bias_set].b[l][j];
      const float* Wl = Ws + l * D * D;
#pragma unroll 2
      for (int k = 0; k < D; ++k) {
        float xv = xs[k * XS_LD + tid];
#pragma unroll
        for (int j4 = 0; j4 < D / 4; ++j4) {
          float4 w = *reinterpret_cast<const float4*>(Wl + k * D + j4 * 4);
          acc[j4 * 4 + 0] = fmaf(xv, w.x, acc[j4 * 4 + 0]);
          acc[j4 * 4 + 1] = fmaf(xv, w.y, acc[j4 * 4 + 1]);
          acc[j4 * 4 + 2] = fmaf(xv, w.z, acc[j4 * 4 + 2]);
          acc[j4 * 4 + 3] = fmaf(xv, w.w, acc[j4 * 4 + 3]);
        }
      }
      const bool last = (MODE == 0) && (l == NL - 1);
      // own column only: no cross-thread hazard
#pragma unroll
      for (int j = 0; j < D; ++j) xs[j * XS_LD + tid] = last ? acc[j] : fmaxf(acc[j], 0.f);
    }
    if (MODE == 0) {
      __syncthreads();
      for (int i = tid; i < SIMT_THREADS * D; i += SIMT_THREADS) {
        int r = i / D, k = i % D;
        if (row0 + r < n_rows) y[(row0 + r) * D + k] = xs[k * XS_LD + r];
      }
    } else {
      float s = c_vote_tail.b4;
#pragma unroll
      for (int k = 0; k < D; ++k) s = fmaf(xs[k * XS_LD + tid], c_vote_tail.w4[k], s);
      if (row0 + tid < n_rows) y[row0 + tid] = s;
    }
  }
}

// ---------------------------------------------------------------------------------
// xV = EV^T . mE : deterministic CSR segment-sum, one warp per vertex row.
// (graphnn.py:156-160 with adjoint_a=True; model.py:76-83)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) simt_segment_sum_kernel(const float* __restrict__ mE,
                                                               const int32_t* __restrict__ vptr,
                                                               const int32_t* __restrict__ vidx, int64_t n_vertices,
                                                               float* __restrict__ xV) {
  int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (v >= n_vertices) return;
  float2 acc = make_float2(0.f, 0.f);
  int beg = vptr[v], end = vptr[v + 1];
  for (int p = beg; p < end; ++p) {
    float2 m = *reinterpret_cast<const float2*>(mE + (int64_t)vidx[p] * D + lane * 2);
    acc.x += m.x;
    acc.y += m.y;
  }
  *reinterpret_cast<float2*>(xV + v * D + lane * 2) = acc;
}

// ---------------------------------------------------------------------------------
// LayerNormBasicLSTMCell step, activation=relu (graphnn.py:107-112,167-170)
//   z = [x, h] . K ; i,j,f,o = split(z) ; LN each ; c' = LN(c*sig(f+1) + sig(i)*relu(j)) ;
//   h' = relu(c') * sig(o)
// GATHER: x[r] = m[src[r]] + m[dst[r]]  (= EV . m, graphnn.py:156-160); else x given.
// One thread per row; K [128][256] fp32 resident in shared memory; in-place state update.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void ln_stats(const float (&v)[D], float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < D; ++j) s += v[j];
  mean = s * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    float t = v[j] - mean;
    q = fmaf(t, t, q);
  }
  rstd = rsqrtf(q * (1.0f / D) + LN_EPS);
}

template <bool GATHER>
__global__ void __launch_bounds__(SIMT_THREADS) simt_lnlstm_kernel(const float* __restrict__ xin,
                                                                   const int32_t* __restrict__ src,
                                                                   const int32_t* __restrict__ dst, int64_t n_rows,
                                                                   const float* __restrict__ K, int cell,
                                                                   float* __restrict__ h, float* __restrict__ c) {
  extern __shared__ float smem[];
  float* Ks = smem;                    // [128][256]
  float* xs = smem + 2 * D * 4 * D;    // [128][XS_LD]  rows 0..63 = x, 64..127 = h
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * D * 4 * D / 4; i += SIMT_THREADS)
    reinterpret_cast<float4*>(Ks)[i] = reinterpret_cast<const float4*>(K)[i];
  const CellLN& ln = c_ln[cell];
  const int64_t n_tiles = (n_rows + SIMT_THREADS - 1) / SIMT_THREADS;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * SIMT_THREADS;
    const int64_t row = row0 + tid;
    const bool valid = row < n_rows;
    __syncthreads();
    if (GATHER) {
      if (valid) {
        const float4* a = reinterpret_cast<const float4*>(xin + (int64_t)src[row] * D);
        const float4* b = reinterpret_cast<const float4*>(xin + (int64_t)dst[row] * D);
#pragma unroll
        for (int k4 = 0; k4 < D / 4; ++k4) {
          float4 u = a[k4], w = b[k4];
          xs[(k4 * 4 + 0) * XS_LD + tid] = u.x + w.x;
          xs[(k4 * 4 + 1) * XS_LD + tid] = u.y + w.y;
          xs[(k4 * 4 + 2) * XS_LD + tid] = u.z + w.z;
          xs[(k4 * 4 + 3) * XS_LD + tid] = u.w + w.w;
        }
      } else {
        for (int k = 0; k < D; ++k) xs[k * XS_LD + tid] = 0.f;
      }
    } else {
      for (int i = tid; i < SIMT_THREADS * D; i += SIMT_THREADS) {
        int r = i / D, k = i % D;
        xs[k * XS_LD + r] = (row0 + r < n_rows) ? xin[(row0 + r) * D + k] : 0.f;
      }
    }
    for (int i = tid; i < SIMT_THREADS * D; i += SIMT_THREADS) {
      int r = i / D, k = i % D;
      xs[(D + k) * XS_LD + r] = (row0 + r < n_rows) ? h[(row0 + r) * D + k] : 0.f;
    }
    __syncthreads();

    float acc[D], keep[D];
    float mean, rstd;
    auto gate_gemm = [&](int g) {
#pragma unroll
      for (int j = 0; j < D; ++j) acc[j] = 0.f;
#pragma unroll 2
      for (int k = 0; k < 2 * D; ++k) {
        float xv = xs[k * XS_LD + tid];
        const float* Kr = Ks + k * 4 * D + g * D;
#pragma unroll
        for (int j4 = 0; j4 < D / 4; ++j4) {
          float4 w = *reinterpret_cast<const float4*>(Kr + j4 * 4);
          acc[j4 * 4 + 0] = fmaf(xv, w.x, acc[j4 * 4 + 0]);
          acc[j4 * 4 + 1] = fmaf(xv, w.y, acc[j4 * 4 + 1]);
          acc[j4 * 4 + 2] = fmaf(xv, w.z, acc[j4 * 4 + 2]);
          acc[j4 * 4 + 3] = fmaf(xv, w.w, acc[j4 * 4 + 3]);
        }
      }
      ln_stats(acc, mean, rstd);
    };
    // input gate
    gate_gemm(0);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float inv = rstd * ln.gamma[0][j];
      keep[j] = 1.0f / (1.0f + expf(-(acc[j] * inv + (ln.beta[0][j] - mean * inv))));
    }
    // transform
    gate_gemm(1);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float inv = rstd * ln.gamma[1][j];
      keep[j] *= fmaxf(acc[j] * inv + (ln.beta[1][j] - mean * inv), 0.f);
    }
    // forget (+ forget_bias after LN) and new cell state
    gate_gemm(2);
    if (valid) {
      const float4* cr = reinterpret_cast<const float4*>(c + row * D);
#pragma unroll
      for (int j4 = 0; j4 < D / 4; ++j4) {
        float4 cv = cr[j4];
        float cc[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int j = j4 * 4 + q;
          float inv = rstd * ln.gamma[2][j];
          float f = acc[j] * inv + (ln.beta[2][j] - mean * inv);
          keep[j] = fmaf(cc[q], 1.0f / (1.0f + expf(-(f + FORGET_BIAS))), keep[j]);
        }
      }
    }
    ln_stats(keep, mean, rstd);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float inv = rstd * ln.gamma[4][j];
      keep[j] = keep[j] * inv + (ln.beta[4][j] - mean * inv);
    }
    if (valid) {
      float4* cw = reinterpret_cast<float4*>(c + row * D);
#pragma unroll
      for (int j4 = 0; j4 < D / 4; ++j4)
        cw[j4] = make_float4(keep[j4 * 4], keep[j4 * 4 + 1], keep[j4 * 4 + 2], keep[j4 * 4 + 3]);
    }
    // output gate and new hidden state
    gate_gemm(3);
    __syncthreads();   // every thread is done reading h rows of xs before h is overwritten
    if (valid) {
      float4* hw = reinterpret_cast<float4*>(h + row * D);
#pragma unroll
      for (int j4 = 0; j4 < D / 4; ++j4) {
        float o4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int j = j4 * 4 + q;
          float inv = rstd * ln.gamma[3][j];
          float o = acc[j] * inv + (ln.beta[3][j] - mean * inv);
          o4[q] = fmaxf(keep[j], 0.f) / (1.0f + expf(-o));
        }
        hw[j4] = make_float4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// logits[k] = mean(vote[off_k : off_k + n_edges[k]]) ; pred = sigmoid(logit)   (model.py:134-147)
// One warp per instance, fixed summation order (deterministic).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) readout_kernel(const float* __restrict__ vote, const int64_t* __restrict__ eoff,
                                                      int n_instances, float* __restrict__ logits,
                                                      float* __restrict__ preds) {
  int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (k >= n_instances) return;
  int64_t beg = eoff[k], end = eoff[k + 1];
  float s = 0.f;
  for (int64_t p = beg + lane; p < end; p += 32) s += vote[p];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    float l = s / (float)(end - beg);
    if (logits) logits[k] = l;
    if (preds) preds[k] = 1.0f / (1.0f + expf(-l));
  }
}

}  // namespace tspgnn
