/* tspgnn.h -- C ABI of the B200-native TSP-GNN message-passing hot path.
 *
 * The reference (machine-reasoning-ufrgs/TSP-GNN) has no FFI: its boundary is the Python
 * surface build_network(d) -> dict + sess.run(fetches, feed_dict) (model.py:9-170,
 * train.py:25-42).  Each entry point below names the piece of that surface it replaces.
 * tsp_gnn_b200/_lib.py binds them with ctypes; INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative TSPGNN_E_* code, with a human-readable message available from
 * tspgnn_last_error() (thread-local).  `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream).  "host" pointers are ordinary (ideally pinned) host
 * memory, "dev" pointers are device memory on the context's device.  Not re-entrant per
 * handle; one handle per host thread / stream (the reference is single-session too).
 */
#ifndef TSPGNN_H_
#define TSPGNN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tspgnn_ctx* tspgnn_handle;

enum {
  TSPGNN_OK = 0,
  TSPGNN_E_INVALID = -1,   /* bad argument / shape (the reference raises Exception, graphnn.py:72-103,185-271) */
  TSPGNN_E_CUDA = -2,      /* CUDA runtime error */
  TSPGNN_E_STATE = -3,     /* call order violated (e.g. forward before plan / set_params) */
  TSPGNN_E_UNSUPPORTED = -4
};

/* Arithmetic of the dense contractions.  State and all LayerNorm / gate math is fp32 in
 * every mode.  */
enum {
  TSPGNN_MODE_SIMT_FP32 = 0,  /* fp32 FMA on CUDA cores (bring-up / cross-check path)            */
  TSPGNN_MODE_TC_BF16X3 = 1,  /* tcgen05, operands split hi+lo bf16, 3 MMAs per product (fp32-parity mode) */
  TSPGNN_MODE_TC_BF16 = 2     /* tcgen05, single bf16 MMA, bf16 embeddings / fp32 accumulate (BASELINE config 3) */
};

const char* tspgnn_last_error(void);
int tspgnn_version(void);

/* Number of float32 values in the parameter blob for embedding size d (115529 for d=64);
 * order = tsp_gnn_b200/params.py:param_spec = the trainable variables of model.py:33-115. */
int64_t tspgnn_param_count(int d);

/* Replaces build_network(d) graph construction + tf.Session() (model.py:9, train.py:202-206).
 * Only d == 64 (the reference default, train.py:108) is implemented here; the Python mirror serves other sizes
 * through the generic building blocks at the end of this header (tsp_gnn_b200/model.py, inference only).  */
int tspgnn_create(int d, int mode, int device, tspgnn_handle* out);
int tspgnn_destroy(tspgnn_handle h);
int tspgnn_get_mode(tspgnn_handle h);

/* Tuning / diagnosis knobs (no equivalent in the reference).
 *   "fused" (default 0): 1 = tspgnn_step runs the persistent fused CTA-pair kernel (tensor-core modes) instead of
 *       the two-kernel sequence (message MLPs, then LSTM cells); both compute the same thing.
 *   "v_pair_weight": relative cost of a vertex tile pair used to split the clusters of the fused kernel.
 *   "train_tc" (default 1): reverse-pass contractions on tcgen05 (tensor-core modes); 0 = fp32 CUDA-core kernels.
 *   "train_graph" (default 1): replay the reverse pass from a CUDA graph once the same tspgnn_backward call
 *       (same plan geometry, buffers and timestep count) has been seen twice in a row.
 *   "act_images" (default 1, bf16x3 mode): tspgnn_train_forward keeps the hidden activations of the edge message MLP
 *       as bf16 hi / lo operand images (75 MB per timestep at the north-star batch) and the reverse pass bulk-copies
 *       them instead of recomputing the three layers per timestep; 0 = recompute.
 *   "d_images" (default 1): inside an MLP's reverse chain d travels between the layer kernels as an operand image;
 *       0 = row-major fp32. */
int tspgnn_set_option(tspgnn_handle h, const char* name, double value);

/* Replaces tf.global_variables_initializer() / Saver.restore (train.py:210, util.py:17):
 * uploads the flat fp32 blob (host pointer) and prepares the on-device operand images. */
int tspgnn_set_params(tspgnn_handle h, const float* host_blob, int64_t n_floats);

/* Replaces feeding the dense EV placeholder plus n_vertices / n_edges (model.py:20-23,
 * instance_loader.py:45-67): edge row e has its two non-zeros in columns edge_src[e] <
 * edge_dst[e] (global vertex ids); instance k owns edge rows [sum n_edges[:k], +n_edges[k])
 * and vertex ids [sum n_vertices[:k], +n_vertices[k]).  Host pointers; copied to the device
 * (inside the call) together with the derived CSR of EV^T.  Workspace is (re)allocated here.  Rows are
 * validated like graphnn.check_run (block-diagonal structure); a batch that fails leaves the handle
 * WITHOUT a plan (TSPGNN_E_INVALID, message names the offending row), the previous one is gone. */
int tspgnn_plan(tspgnn_handle h, int n_instances, const int32_t* n_vertices, const int32_t* n_edges,
                const int32_t* edge_src, const int32_t* edge_dst);

/* Replaces sess.run(predictions, feed_dict={W, C, time_steps}) (train.py:25-42): host W,C
 * [sumE] fp32 are copied in, the forward pass runs, logits/predictions [n_instances] fp32
 * are copied back to host; synchronises `stream` before returning.  Either output may be NULL. */
int tspgnn_forward_host(tspgnn_handle h, const float* W, const float* C, int time_steps,
                        float* logits, float* predictions, void* stream);

/* Same with W, C, logits, predictions in device memory; asynchronous on `stream`. */
int tspgnn_forward_device(tspgnn_handle h, const float* dW, const float* dC, int time_steps,
                          float* d_logits, float* d_predictions, void* stream);

/* The three phases of the forward pass, separately (device pointers, asynchronous):
 *   init_embeddings : model.py:33-51 + graphnn.py:134-139  (E0 = E_init_MLP([W,C]), V0, c = 0)
 *   step            : n_steps iterations of while_body, graphnn.py:142-173 -- the unit the
 *                     "message-passing timesteps/s" metric counts
 *   readout         : model.py:107-147 (E_vote MLP, per-instance mean, sigmoid)            */
int tspgnn_init_embeddings(tspgnn_handle h, const float* dW, const float* dC, void* stream);
int tspgnn_step(tspgnn_handle h, int n_steps, void* stream);
int tspgnn_readout(tspgnn_handle h, float* d_logits, float* d_predictions, void* stream);

/* Replaces fetching last_states (model.py:123): row-major fp32 [sumV,64] / [sumE,64]
 * device buffers; any pointer may be NULL.  Asynchronous on `stream`. */
int tspgnn_get_states(tspgnn_handle h, float* dVh, float* dVc, float* dEh, float* dEc, void* stream);
/* LSTM_initial_states / initial_embeddings of GraphNN.__call__ (graphnn.py:128,134-139):
 * overwrite the recurrent state from row-major fp32 device buffers (NULL = leave as is). */
int tspgnn_set_states(tspgnn_handle h, const float* dVh, const float* dVc, const float* dEh,
                      const float* dEc, void* stream);

/* ---- training step: replaces sess.run([train_step, loss, ...]) (train.py:35-42, model.py:157-167) ----
 *
 * tspgnn_train_forward : the forward pass of tspgnn_forward_device that also keeps, per timestep,
 *                        the recurrent state the reverse pass needs (tf.while_loop's loop stacks).
 * tspgnn_backward      : loss = sum_k xent(logit_k, route_exists_k) / global_batch (model.py:157) and
 *                        d loss / d every trainable variable (tf.gradients, model.py:166) into a flat
 *                        device blob laid out like the parameter blob.  global_batch is the divisor
 *                        of reduce_mean: the number of instances of the WHOLE batch when instances are
 *                        sharded over ranks (then the sum of the ranks' blobs is the batch gradient);
 *                        <= 0 means this handle's own batch.  d_grads NULL = the handle's own buffer
 *                        (tspgnn_grad_buffer).  d_loss (device, 1 float) may be NULL.
 * tspgnn_apply_gradients: model.py:160-167: adds l2norm_scaling * var (gradient of the L2 term),
 *                        clip_by_global_norm(0.65), one Adam step (lr 2e-5, TF defaults), then
 *                        refreshes every derived operand image.  Synchronises `stream`.
 * tspgnn_train_step_host: all of the above with host buffers; loss / logits / predictions are the
 *                        values computed with the variables BEFORE the update, like the reference's
 *                        single sess.run. */
int tspgnn_train_forward(tspgnn_handle h, const float* dW, const float* dC, int time_steps, float* d_logits,
                         float* d_predictions, void* stream);
int tspgnn_backward(tspgnn_handle h, const float* d_route_exists, int global_batch, float* d_grads, float* d_loss,
                    void* stream);
int tspgnn_apply_gradients(tspgnn_handle h, float* d_grads, float* host_global_norm, void* stream);
int tspgnn_train_step_host(tspgnn_handle h, const float* W, const float* C, const float* route_exists, int time_steps,
                           float* loss, float* logits, float* predictions, void* stream);
float* tspgnn_grad_buffer(tspgnn_handle h);

/* Current value of every trainable variable (host blob, tspgnn_param_count floats): what
 * util.save_weights persists (util.py:24-37). */
int tspgnn_get_params(tspgnn_handle h, float* host_blob, int64_t n_floats);

/* Hyper-parameters of model.py:13-15 and tf.train.AdamOptimizer (defaults: 2e-5, 1e-10, 0.65,
 * 0.9, 0.999, 1e-8), and the Adam slots tf.train.Saver stores next to the variables (util.py:35). */
int tspgnn_set_hyper(tspgnn_handle h, float learning_rate, float l2norm_scaling, float clip_norm, float beta1,
                     float beta2, float epsilon);
int tspgnn_get_optimizer_state(tspgnn_handle h, float* host_m, float* host_v, int64_t* step, int64_t n_floats);
int tspgnn_set_optimizer_state(tspgnn_handle h, const float* host_m, const float* host_v, int64_t step,
                               int64_t n_floats);

/* Sizes of the current plan. */
int64_t tspgnn_sum_edges(tspgnn_handle h);
int64_t tspgnn_sum_vertices(tspgnn_handle h);

/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int64_t tspgnn_launch_count(tspgnn_handle h);

/* Measurement hook for bench.py's roofline line: launches ONE kernel of the timestep
 * `iters` times on `stream`, each launch bracketed by CUDA events, and returns the mean
 * duration in milliseconds.  which: 0 = LayerNorm-LSTM kernel (K1), 1 = message-MLP kernel
 * (K2), 2 = the fused timestep kernel (cells + messages of the new state, what tspgnn_step launches).  Tensor-core modes only.  The recurrent state keeps evolving while this runs. */
int tspgnn_time_kernel(tspgnn_handle h, int which, int iters, float* mean_ms, void* stream);

/* Development aid: one launch of K1 (which = 0), K2 (which = 1) or the fused kernel (which = 2) with a clock64() trace of the
 * warp roles, copied to out_host[cta][role][tile][event] (148*4*64*8 int64). */
int tspgnn_debug_timeline(tspgnn_handle h, int which, long long* out_host, int64_t n_int64, void* stream);

/* ---- generic building blocks (stateless; fp32 row-major device pointers on `device`, asynchronous on `stream`) ----
 * They execute the parts of the reference's public surface that the fused TSP kernels do not cover:
 * stand-alone Mlp evaluation and GraphNN topologies other than build_network's (a transfer function, matrix-only
 * inputs, several update terms per variable; graphnn.py:142-173).  Activation codes: 0 none, 1 relu, 2 tanh, 3 sigmoid.
 *
 * tspgnn_dense_forward  : one tf.layers.Dense call (mlp.py:39-52,57-63): Y[rows,out] = act(X[rows,in] . W[in,out] + b);
 *                         dB may be NULL (use_bias=False).
 * tspgnn_matmul_coo     : tf.matmul(adjacency, y, adjoint_a=transpose) (graphnn.py:155-161) for a matrix given by its
 *                         stored entries (row, col, value; d_val NULL = all ones): dOut[out_rows, d] is overwritten.
 * tspgnn_lnlstm_forward : one LayerNormBasicLSTMCell call (graphnn.py:107-112,167-170) with input width in_dim and
 *                         `units` units: dXH = [x, h] concatenated [rows, in_dim + units], kernel [(in_dim+units), 4 units],
 *                         gamma / beta [5][units] in gate order input, transform, forget, output, state;
 *                         d_scratch holds rows * 4 * units floats. */
int tspgnn_dense_forward(int device, const float* dX, int64_t rows, int in_dim, const float* dW, const float* dB,
                         int out_dim, int activation, float* dY, void* stream);
int tspgnn_matmul_coo(int device, const int32_t* d_row, const int32_t* d_col, const float* d_val, int64_t nnz,
                      int transpose, const float* dY, int d, int64_t out_rows, float* dOut, void* stream);
int tspgnn_lnlstm_forward(int device, const float* dXH, int in_dim, const float* dC, int64_t rows, int units,
                          const float* dKernel, const float* dGamma, const float* dBeta, int activation,
                          float forget_bias, float* dC_out, float* dH_out, float* d_scratch, void* stream);

/* Host helper: dense EV (row-major [rows, cols], float64 or float32 by elem_size 8/4) ->
 * edge_src/edge_dst (instance_loader.py:63-66 layout).  Returns TSPGNN_E_INVALID if a row
 * does not have exactly two non-zeros. */
int tspgnn_dense_ev_to_coo(const void* EV, int elem_size, int64_t rows, int64_t cols,
                           int32_t* edge_src, int32_t* edge_dst);

#ifdef __cplusplus
}
#endif
#endif /* TSPGNN_H_ */
