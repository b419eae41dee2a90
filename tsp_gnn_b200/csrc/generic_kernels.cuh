// Generic (any size, fp32) building blocks behind the reference's public pieces that the fused
// TSP kernels do not cover: a stand-alone dense layer (Mlp.__call__, mlp.py:57-63), the product with
// an arbitrary adjacency matrix in coordinate form (graphnn.py:155-161) and a LayerNormBasicLSTMCell
// of any input width / unit count (graphnn.py:107-112,167-170).  GraphNN topologies other than the
// TSP wiring (a transfer function, matrix-only inputs, several update terms per variable) are
// executed from these; they are not the hot path, so the kernels are plain CUDA-core code.
#pragma once
#include "common.cuh"

namespace tspgnn {

// activations by code: 0 none, 1 relu, 2 tanh, 3 sigmoid
__device__ __forceinline__ float generic_act(float x, int act) {
  if (act == 1) return fmaxf(x, 0.f);
  if (act == 2) return tanhf(x);
  if (act == 3) return 1.0f / (1.0f + expf(-x));
  return x;
}

// Y[rows, N] = act(X[rows, K] . W[K, N] + b)      (tf.layers.Dense: kernel [in, out])
// 64 x 64 output tile per CTA of 256 threads, 4 x 4 register tile per thread, K in chunks of 16.
__global__ void __launch_bounds__(256) generic_dense_kernel(const float* __restrict__ X, int64_t rows, int K,
                                                            const float* __restrict__ W, const float* __restrict__ b,
                                                            int N, int act, float* __restrict__ Y) {
  __shared__ float xs[16][64 + 1];
  __shared__ float ws[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int col0 = blockIdx.y * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int kk = i & 15, r = i >> 4;
      const int64_t gr = row0 + r;
      xs[kk][r] = (gr < rows && k0 + kk < K) ? X[gr * K + k0 + kk] : 0.f;
    }
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int c = i & 63, kk = i >> 6;
      ws[kk][c] = (k0 + kk < K && col0 + c < N) ? W[static_cast<int64_t>(k0 + kk) * N + col0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float xv[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = xs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty * 4 + i;
    if (gr >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = col0 + tx * 4 + j;
      if (gc < N) Y[gr * N + gc] = generic_act(acc[i][j] + (b ? b[gc] : 0.f), act);
    }
  }
}

// out[r_out[k], :] += val[k] * y[r_in[k], :]  for every stored entry k of a matrix in coordinate form
// (out must be zero before the launch)
__global__ void __launch_bounds__(256) generic_coo_matmul_kernel(const int32_t* __restrict__ r_out,
                                                                 const int32_t* __restrict__ r_in,
                                                                 const float* __restrict__ val, int64_t nnz, int d,
                                                                 const float* __restrict__ y, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nnz * d) return;
  const int64_t k = i / d;
  const int j = static_cast<int>(i % d);
  const float v = val ? val[k] : 1.0f;
  atomicAdd(out + static_cast<int64_t>(r_out[k]) * d + j, v * y[static_cast<int64_t>(r_in[k]) * d + j]);
}

__device__ __forceinline__ float generic_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Gate math of LayerNormBasicLSTMCell for any unit count: Z[rows, 4u] = [x, h] . kernel (no bias) comes
// from generic_dense_kernel; one warp per row.  Gate order input, transform, forget, output; forget
// bias added after the gate's LayerNorm; the LayerNorm'd new c is what is carried (TF 1.x contrib).
__global__ void __launch_bounds__(256) generic_lnlstm_gates_kernel(const float* __restrict__ Z, const float* __restrict__ c_in,
                                                                   int64_t rows, int u, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, int act,
                                                                   float forget_bias, float* __restrict__ c_out,
                                                                   float* __restrict__ h_out) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* z = Z + row * 4 * u;
  float mean[4], rstd[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float s = 0.f;
    for (int j = lane; j < u; j += 32) s += z[g * u + j];
    const float m = generic_warp_sum(s) / u;
    float q = 0.f;
    for (int j = lane; j < u; j += 32) {
      const float t = z[g * u + j] - m;
      q = fmaf(t, t, q);
    }
    mean[g] = m;
    rstd[g] = rsqrtf(generic_warp_sum(q) / u + LN_EPS);
  }
  auto ln = [&](int g, int j) { return (z[g * u + j] - mean[g]) * rstd[g] * gamma[g * u + j] + beta[g * u + j]; };
  // new cell state before its LayerNorm
  float s = 0.f;
  for (int j = lane; j < u; j += 32) {
    const float i = ln(0, j), jj = ln(1, j), f = ln(2, j);
    const float cn = c_in[row * u + j] * (1.0f / (1.0f + expf(-(f + forget_bias)))) +
                     (1.0f / (1.0f + expf(-i))) * generic_act(jj, act);
    c_out[row * u + j] = cn;       // parked; normalised in place below
    s += cn;
  }
  const float cm = generic_warp_sum(s) / u;
  float q = 0.f;
  for (int j = lane; j < u; j += 32) {
    const float t = c_out[row * u + j] - cm;
    q = fmaf(t, t, q);
  }
  const float crs = rsqrtf(generic_warp_sum(q) / u + LN_EPS);
  for (int j = lane; j < u; j += 32) {
    const float cn = (c_out[row * u + j] - cm) * crs * gamma[4 * u + j] + beta[4 * u + j];
    const float o = ln(3, j);
    c_out[row * u + j] = cn;
    h_out[row * u + j] = generic_act(cn, act) * (1.0f / (1.0f + expf(-o)));
  }
}

}  // namespace tspgnn
