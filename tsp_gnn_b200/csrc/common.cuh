// Shared constants and small helpers for the TSP-GNN hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tspgnn {

constexpr int D = 64;             // embedding size (train.py:108 default; the only size built)
constexpr int TILE_ROWS = 128;    // rows (edges or vertices) per tile = UMMA M = TMEM lanes
constexpr float LN_EPS = 1e-12f;  // tf.contrib.layers.layer_norm variance_epsilon
constexpr float FORGET_BIAS = 1.0f;

// Layer-norm parameters of one LayerNormBasicLSTMCell, gate order input, transform,
// forget, output, state (graphnn.py:107-112).  Lives in __constant__ memory: every
// thread of a warp touches the same column at the same time, so reads are broadcasts.
struct CellLN {
  float gamma[5][D];
  float beta[5][D];
};

// Biases of a 4-layer message MLP (graphnn.py:114-125); layer 4 of the vote MLP is the
// 64->1 column stored in w4 / b4.
struct MlpBias {
  float b[4][D];
};

struct VoteTail {
  float w4[D];
  float b4;
};

struct EInit {      // model.py:33-43, sizes 2->8->16->32->64
  float w1[2][8], b1[8];
  float w2[8][16], b2[16];
  float w3[16][32], b3[32];
  float w4[32][64], b4[64];
};

}  // namespace tspgnn
