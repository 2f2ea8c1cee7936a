/*
 * magnet_b200 — C ABI of the B200-native MAgNet hot path (libmagnet_b200.so).
 *
 * This header is the drop-in boundary.  Every entry point names the reference interface
 * (jaggbow/magnet, file:line) it replaces.  Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless marked host;
 *   - the CALLER owns every buffer, including workspaces (size them with the *_workspace
 *     queries); the library never allocates or frees device memory;
 *   - every call takes an explicit cudaStream_t (pass as void*), assumes the caller has set
 *     the device, is re-entrant, and never synchronises the host unless stated;
 *   - return 0 on success, <0 on error (-1 bad argument, -2 workspace too small, -3 capacity
 *     overflow, -4 CUDA error); mgb_last_error() returns a thread-local message;
 *   - features are fp32 row-major; public indices are int64 exactly as the reference's
 *     (edge_index, assign_index); private plans are int32.
 * Hidden width is fixed at 128 (hidden_features / latent_dim / n_chan of every reference config).
 */
#ifndef MAGNET_B200_H
#define MAGNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_ABI_VERSION 1

int mgb_abi_version(void);
const char* mgb_last_error(void);
/* kernels launched by this library in this process so far (bench.py reports the per-step delta) */
int64_t mgb_launch_count(void);
/* Optional per-kernel device timing (CUDA events on the launching stream) for bench.py's roofline.
 * kernel ids: 0 GNN edge fwd, 1 GNN edge bwd, 2 node GEMM, 3 weight-grad, 4 graph build,
 * 5 InteractionNetwork edge fwd, 6 InteractionNetwork edge bwd, 7 INR decode.
 * total_ms / count are HOST pointers; collect synchronises on the recorded events. */
int mgb_profile_enable(int on);
int mgb_profile_collect(int kernel_id, double* total_ms, int64_t* count);

/* ---------------------------------------------------------------------------------------------
 * Graph construction.
 * Replaces torch_geometric.nn.radius_graph as called at models/mpnn_2d.py:245, models/mpnn.py:245
 * (loop=False) and models/magnet_gnn.py:293 (loop=True); torch_cluster CUDA ordering, default
 * max_num_neighbors = 32, bit-exact including the index-order truncation.
 *
 * Phase 1 (no host sync): neighbour lists.  cap = max_num_neighbors + (loop ? 0 : 1).
 *   pos [n, d] fp32 (d = 1 or 2), ptr [n_samples+1] int64 node offsets of the samples,
 *   nbr [n, cap] int32 (out), deg [n] int32 (out), rowptr [n+1] int32 (out; rowptr[n] = E).
 * Phase 2: edge_index [2, E] int64 in the reference's order (sorted by centre, then neighbour);
 *   centre_row = 1 gives PyG's [neighbour; centre] (mpnn*), centre_row = 0 gives MAgNetGNN's
 *   swapped [centre; neighbour] (models/magnet_gnn.py:294-296).  col [E] int32 (optional) = the
 *   neighbour of every edge.
 * ------------------------------------------------------------------------------------------- */
size_t mgb_radius_graph_workspace(int64_t n, int n_samples);
int mgb_radius_graph_search(const float* pos, int64_t n, int d, const int64_t* ptr, int n_samples, double r,
                            int max_num_neighbors, int loop, int32_t* nbr, int32_t* deg, int32_t* rowptr,
                            void* workspace, size_t workspace_bytes, void* stream);
int mgb_radius_graph_emit(const int32_t* nbr, const int32_t* rowptr, int64_t n, int cap, int centre_row,
                          int64_t n_edges, int64_t* edge_index, int32_t* col, void* stream);

/* Replaces torch_geometric.nn.knn as called at models/magnet_gnn.py:247.
 *   x [nx, d] (searched set), y [ny, d] (queries), ptr_x/ptr_y [n_samples+1] int64, k <= 64.
 *   out_idx [ny, k] int64: ascending distance, ties -> lower index, -1 where the sample has < k rows.
 *   out_dist [ny, k] fp32 (optional, may be NULL). */
size_t mgb_knn_workspace(int64_t nx, int n_samples);
int mgb_knn(const float* x, int64_t nx, const float* y, int64_t ny, int d, const int64_t* ptr_x, const int64_t* ptr_y,
            int n_samples, int k, int64_t* out_idx, float* out_dist, void* workspace, size_t workspace_bytes,
            void* stream);

/* Aggregation plan for an arbitrary edge_index (what MessagePassing.propagate + scatter(reduce='mean')
 * do implicitly, models/mpnn_2d.py:46,69; models/magnet_gnn.py:54,76): stable sort of the edges by
 * the aggregation endpoint agg = edge_index[1], plus the transposed plan by other = edge_index[0].
 *   rowptr [n_nodes+1], perm [E] (COO edge id at each position), dst [E], src [E],
 *   rowptr_t [n_nodes+1], pos_t [E] (positions grouped by source) — all int32;
 *   bad_flag [1] int32 device: set to 1 when an index is outside [0, n_nodes). */
size_t mgb_csr_plan_workspace(int64_t n_edges);
int mgb_csr_plan(const int64_t* agg, const int64_t* other, int64_t n_edges, int64_t n_nodes, int32_t* rowptr,
                 int32_t* perm, int32_t* dst, int32_t* src, int32_t* rowptr_t, int32_t* pos_t, int32_t* bad_flag,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MP-PDE message-passing layer.  Replaces GNN_Layer.forward (models/mpnn_2d.py:65-71; message
 * :73-79, update :81-90, mean aggregation :46, InstanceNorm :63,70) and its autograd backward.
 *   x [N,128], u [N,tw], pos [N,dp], var [N,nv];  dp = 2 (mpnn_2d) or 1 (mpnn), nv = 1.
 *   plan from mgb_csr_plan (or mgb_radius_graph_*: for PyG-ordered graphs rowptr/col ARE the plan);
 *   gptr [G+1] int64 node offsets of the graphs (the `batch` vector as offsets).
 *   Weights in PyTorch layout: W1 [128, 256+tw+dp+nv], W2 [128,128], W3 [128, 256+nv], W4 [128,128].
 *   packed: mgb_gnn_layer_packed_floats() floats, filled by mgb_gnn_layer_pack (re-run when the
 *   parameters change).
 *   Saved for backward: pq [N,256], agg [N,128], y1_pre [N,128], y2_pre [N,128], rstd [G,128].
 *   precision selects the arithmetic of the per-edge contraction (the 128x128 message_net_2 product and
 *   its gradients):  0 = fp32 FFMA;  1 = tcgen05 tensor cores, bf16 hi/lo split of both operands with
 *   fp32 accumulation in TMEM (error ~2^-17, meets the 1e-5 fp32 contract);  2 = tcgen05 plain bf16
 *   operands + tanh.approx Swish (1e-2 contract).
 * ------------------------------------------------------------------------------------------- */
size_t mgb_gnn_layer_packed_floats(int tw, int dp, int nv);
int mgb_gnn_layer_pack(const float* W1, const float* b1, const float* W2, const float* W3, const float* W4, int tw,
                       int dp, int nv, float* packed, void* stream);
size_t mgb_gnn_layer_fwd_workspace(int64_t n_nodes, int64_t n_edges, int n_graphs, int max_nodes_per_graph);
int mgb_gnn_layer_fwd(int64_t n_nodes, int64_t n_edges, int tw, int dp, int nv, int n_graphs, int max_nodes_per_graph,
                      const float* x, const float* u, const float* pos, const float* var, const int32_t* rowptr,
                      const int32_t* dst, const int32_t* src, const int64_t* gptr, const float* packed,
                      const float* b2, const float* b3, const float* b4, float* y, float* pq, float* agg,
                      float* y1_pre, float* y2_pre, float* rstd, int precision, void* workspace,
                      size_t workspace_bytes, void* stream);
size_t mgb_gnn_layer_bwd_workspace(int64_t n_nodes, int64_t n_edges, int tw, int dp, int nv, int n_graphs,
                                   int max_nodes_per_graph);
/* du / dpos / dvar may be NULL.  accumulate_params != 0 adds into the d* parameter buffers. */
int mgb_gnn_layer_bwd(int64_t n_nodes, int64_t n_edges, int tw, int dp, int nv, int n_graphs, int max_nodes_per_graph,
                      const float* dy, const float* x, const float* u, const float* pos, const float* var,
                      const float* y, const float* pq, const float* agg, const float* y1_pre, const float* y2_pre,
                      const float* rstd, const int32_t* rowptr, const int32_t* dst, const int32_t* src,
                      const int32_t* rowptr_t, const int32_t* pos_t, const int64_t* gptr, const float* packed,
                      const float* W2, const float* b2, const float* W3, const float* W4, float* dx, float* du,
                      float* dpos, float* dvar, float* dW1, float* db1, float* dW2, float* db2, float* dW3,
                      float* db3, float* dW4, float* db4, int accumulate_params, int precision, void* workspace,
                      size_t workspace_bytes, void* stream);

/* GNN_Layer.update (models/mpnn_2d.py:81-90; models/mpnn.py:81-90) without the InstanceNorm, as ONE launch (what
 * mgb_gnn_layer_fwd runs between the edge kernel and the InstanceNorm; exported for tests and for callers with their own
 * aggregation):
 *   y1_pre = [x, agg, var] W3^T + b3;   y2_pre = Swish(y1_pre) W4^T + b4;   out = x + Swish(y2_pre)
 * x, agg [N,128]; var [N,nv], nv <= 4; W3 [128, 256+nv], W4 [128,128] in PyTorch layout.  packed:
 * mgb_gnn_node_update_packed_floats() floats filled by mgb_gnn_node_update_pack (tensor-memory operand images of W3 / W4 and
 * the var columns of W3; re-run when the parameters change).  precision 1 (bf16 hi/lo split, fp32 contract) or 2 (bf16). */
size_t mgb_gnn_node_update_packed_floats(void);
int mgb_gnn_node_update_pack(const float* W3, const float* W4, int nv, float* packed, void* stream);
int mgb_gnn_node_update_fwd(const float* x, const float* agg, const float* var, int nv, int64_t n_nodes, const float* packed,
                            const float* b3, const float* b4, float* y1_pre, float* y2_pre, float* out, int precision,
                            void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row-wise dense stages (nn.Linear + activation, nn.LayerNorm) used by embedding_mlp
 * (models/mpnn_2d.py:130-135), MLP (models/backbones/mlp.py:9-28), Encoder/Decoder/projector
 * (models/magnet_gnn.py:11-42,119-137,194-197).  act: 0 none, 1 ReLU, 2 Swish.
 *   y = act(x W^T + b) (+ residual);  W [out, in] PyTorch layout; wt = W^T [in, out] (mgb_transpose).
 * ------------------------------------------------------------------------------------------- */
int mgb_transpose(const float* in, int rows, int cols, float* out, void* stream);
int mgb_linear_fwd(const float* x, int64_t rows, int in_features, int out_features, const float* wt, const float* bias,
                   int act, const float* residual, float* y, float* y_pre, void* stream);
/* The same Linear on the tensor cores (tcgen05), for in_features 128 or 256 and out_features <= 256
 * (mgb_linear_tc_packed_floats returns 0 for anything else).  precision: 1 bf16 hi/lo split (16 significant bits),
 * 2 plain bf16 (1e-2 contract), 3 fp16 hi/lo split (22 significant bits; inputs and weights must be O(1), |x| < 65504).
 * packed = swizzled 16-bit images of W [out, in] (row stride ldw) in the format of `precision`, built once per weight
 * version by mgb_linear_tc_pack; for the shapes mgb_linear_tc_bwd covers and precision 1 / 2 the block also holds the blocks of
 * W^T in tensor-memory operand order (the data gradient runs with its weights resident in tensor memory). */
/* fp16 range guard of the precision-3 kernels (mgb_linear_tc_fwd*, mgb_mlp_chain_fwd, mgb_inr_decode_fused, the InteractionNetwork
 * kernels): a kernel that meets |x| >= 32768 raises a host-mapped flag; the next precision-3 call fails with an error, and
 * mgb_f16_range_check() returns 1 (clearing the flag) so that a caller can test at its own synchronisation points — synchronise
 * the stream first for an answer that covers everything issued so far. */
int mgb_f16_range_check(void);
size_t mgb_linear_tc_packed_floats(int in_features, int out_features);
int mgb_linear_tc_pack(const float* W, int ldw, int in_features, int out_features, int precision, float* packed,
                       void* stream);
int mgb_linear_tc_fwd(const float* x, int64_t rows, int in_features, int out_features, const float* packed,
                      const float* bias, int act, const float* residual, float* y, float* y_pre, int precision,
                      void* stream);
/* The same Linear applied to cat([x0, x1], dim = 1) of two 128-column tensors (row strides ld0, ld1) without materialising the
 * concatenation: the node function of InteractionNetwork reads cat([aggregated, x]) (models/magnet_gnn.py:84-86).  Inference. */
int mgb_linear_tc_fwd2(const float* x0, int ld0, const float* x1, int ld1, int64_t rows, int out_features, const float* packed,
                       const float* bias, int act, float* y, int precision, void* stream);
/* A whole MLP of models/backbones/mlp.py:9-28 behind its first Linear — n_layers Linear(128, 128) + act, the last one
 * Linear(128, n_out <= 128) without activation — in one launch, forward only (inference / rollout): activations stay in
 * shared memory between the layers, arithmetic as mgb_linear_tc_fwd with precision 3.  x [rows, >=128] fp32 (row stride
 * ldx, optional ReLU applied on load: in_act = 1), y [rows, n_out] (row stride ldy).  packed: filled layer by layer with
 * mgb_mlp_chain_pack_layer (W_l [out, 128] row stride ldw, bias_l [out] or NULL). */
size_t mgb_mlp_chain_packed_floats(int n_layers);
int mgb_mlp_chain_pack_layer(const float* W, int ldw, int out_features, const float* bias, int layer, int n_layers,
                             float* packed, void* stream);
int mgb_mlp_chain_fwd(const float* x, int ldx, int64_t rows, int n_layers, const float* packed, int act, int in_act,
                      int n_out, float* y, int ldy, void* stream);
/* The same backward on the tensor cores, for out_features == 128 and in_features 128 or 256 (the workspace query returns 0
 * for anything else): dx through the weight images of mgb_linear_tc_pack read MN-major, dW and db in one split-K launch with
 * TMEM-resident accumulators.  precision 1 (bf16 hi/lo split; gradients span too many binades for the fp16 split) or 2 (bf16). */
size_t mgb_linear_tc_bwd_workspace(int64_t rows, int in_features, int out_features);
int mgb_linear_tc_bwd(const float* dy, const float* y_pre, int act, const float* x, int64_t rows, int in_features, int out_features,
                      const float* packed, float* dx, float* dw, float* db, int accumulate_params, int precision, void* workspace,
                      size_t workspace_bytes, void* stream);
size_t mgb_linear_bwd_workspace(int64_t rows, int in_features, int out_features);
/* dx = (dy * act'(y_pre)) W;  dW (+)= (dy * act'(y_pre))^T x;  db (+)= colsum.  dx may be NULL. */
int mgb_linear_bwd(const float* dy, const float* y_pre, int act, const float* x, int64_t rows, int in_features,
                   int out_features, const float* w, float* dx, float* dw, float* db, int accumulate_params,
                   void* workspace, size_t workspace_bytes, void* stream);
/* ---------------------------------------------------------------------------------------------
 * Flat-buffer Adam: torch.optim.Adam(lr, weight_decay) as configured by models/magnet_gnn.py:378-386 and
 * models/mpnn_2d.py:205-213, one launch for all parameters of a model (param / grad / moments are flat fp32 buffers of
 * n elements, 16-byte aligned).  step >= 1 counts this update; grad_scale multiplies the gradient first (1/world
 * after a sum all-reduce); the StepLR schedule is the caller's `lr`.
 * ------------------------------------------------------------------------------------------- */
int mgb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int64_t step, double grad_scale, void* stream);
int mgb_layernorm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int cols, float* y,
                      float* stats /* [rows,2] mean,rstd */, void* stream);
/* y = LayerNorm(x) + residual in one pass (the `x_new + x` of InteractionNetwork.forward, models/magnet_gnn.py:88).  Inference. */
int mgb_layernorm_residual_fwd(const float* x, const float* gamma, const float* beta, const float* residual, int64_t rows, int cols,
                               float* y, void* stream);
size_t mgb_layernorm_bwd_workspace(int64_t rows, int cols);
int mgb_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* stats, int64_t rows, int cols,
                      float* dx, float* dgamma, float* dbeta, int accumulate_params, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Temporal-bundling decoder + Euler update of MP-PDE (models/mpnn_2d.py:138-162 `output_mlp`, :196-200; models/mpnn.py:139-162,
 * :196-200) in one launch per direction:
 *   out[n, j] = u[n, u_col] + (j + 1) dt * Conv1d(8 -> 1, k2)(act(Conv1d(1 -> 8, k1, stride1)(h[n, None, :])))[j],  j < time_window
 * h [N, hidden = 128]; w1 = output_mlp[0].weight [8,1,k1], w2 = output_mlp[-1].weight [1,8,k2]; act: 0 none (the 1-D tw = 10
 * decoder has no Swish, models/mpnn.py:139-142), 2 Swish; dt: DEVICE scalar (the reference derives it from the batch, :317 — no
 * host sync).  bwd: dh [N,128], du[:, u_col] (row stride lddu; NULL to skip), dw1, db1, dw2, db2 overwritten; fixed-order sums. */
int mgb_bundling_decoder_fwd(const float* h, int64_t n, int hidden, const float* u, int ldu, int u_col, const float* w1, const float* b1,
                             int k1, int stride1, const float* w2, const float* b2, int k2, int time_window, int act, const float* dt,
                             float* out, void* stream);
size_t mgb_bundling_decoder_bwd_workspace(int64_t n);
int mgb_bundling_decoder_bwd(const float* h, int64_t n, int hidden, const float* u, int ldu, int u_col, const float* w1, const float* b1,
                             int k1, int stride1, const float* w2, const float* b2, int k2, int time_window, int act, const float* dt,
                             const float* dout, float* dh, float* du, int lddu, float* dw1, float* db1, float* dw2, float* db2,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MAgNet[GNN] InteractionNetwork glue (models/magnet_gnn.py:70-90).  The first Linear of edge_fn is
 * factorised, W [x_i, x_j, e] = P[i] + Q[j] + R[e] with i = edge_index[1] (the aggregation endpoint),
 * j = edge_index[0]; per-edge tensors are [E,128] in the caller's COO order.
 *   mgb_edge_combine_fwd:  out[e] = act(p[i_e] + q[j_e] + r[e]), act 0 none / 1 ReLU; edge_index int64 [2,E].
 *   mgb_relu_mask:         dz = dout * (out > 0).
 *   mgb_segment_sum_rows:  out[n] = scale * sum_{q in [rowptr[n], rowptr[n+1])} rows[idx ? idx[q] : q]; scale = 1/max(count,1)
 *                          when mean != 0 (scatter(reduce='mean'), aggr='mean' :54), fixed order, no atomics.
 *   mgb_gather_rows:       out[e] = rows[index[e]] * (rowptr ? 1/max(count(index[e]),1) : 1)  (backward of the mean).
 * ------------------------------------------------------------------------------------------- */
int mgb_edge_combine_fwd(const float* p, const float* q, const float* r, const int64_t* edge_index, int64_t n_edges, int act,
                         float* out, void* stream);
int mgb_relu_mask(const float* dout, const float* out, int64_t n, float* dz, void* stream);
int mgb_segment_sum_rows(const float* rows, int cols, const int32_t* rowptr, const int32_t* idx, int64_t n_nodes, int mean,
                         float* out, int ld_out, void* stream);
int mgb_gather_rows(const float* rows, const int64_t* index, const int32_t* rowptr, int64_t n_edges, float* out, void* stream);
/* Node and edge features of MAgNetGNN._build_graph (models/magnet_gnn.py:298-308) in one launch:
 *   node_features [N, C+d+1] = [u, x, t_last[row % n_samples]] (the time column is TILED over the rows),
 *   edge_features [E, C+d]   = [u[s] - u[r], x[s] - x[r]] with s = edge_index[0], r = edge_index[1] (int64 [2,E]).  Inference. */
int mgb_magnet_features(const float* u, int n_chan, const float* x, int d, const float* t_last, int n_samples, int64_t n_nodes,
                        const int64_t* edge_index, int64_t n_edges, float* node_features, float* edge_features, void* stream);

/* Fused edge function of InteractionNetwork (models/magnet_gnn.py:79-82 message, :54 aggr='mean'), forward:
 *   agg[i] = mean_{e: edge_index[1][e] = i} LayerNorm(MLP5(cat[x_i, x_j, e_scale * e_features[e]]))
 * in ONE launch: e_features rows are gathered through the aggregation plan, the five 128x128 contractions of edge_fn run
 * as tcgen05 tiles whose activations stay in shared memory, LayerNorm and the per-destination mean (fixed order, no
 * atomics) happen in the last epilogue.  Nothing of size [E,128] is written.
 *   pq [N,256] = x [W0[:, :128] | W0[:, 128:256]]^T + [b0 | 0]  (first Linear factorised per node, mgb_linear_tc_fwd);
 *   perm / rowptr / dst / src: plan of mgb_csr_plan (perm NULL: e_features already in aggregation order);
 *   e_scale: the reference doubles e_features every layer and never updates them - layer l sees 2^l e_0;
 *   packed: mgb_in_edge_packed_floats() floats; layer 0 = W0[:, 256:384] (ldw 384, bias NULL), layers 1..4 = the other
 *   Linears of edge_fn with their biases, then the LayerNorm affine.  precision: 2 = bf16 operands (1e-2 contract),
 *   otherwise fp16 hi/lo split (1e-5 contract).  agg [N,128] is written for every node (0 where no edge arrives). */
size_t mgb_in_edge_packed_floats(void);
int mgb_in_edge_pack_layer(const float* W, int ldw, const float* bias, int layer, int precision, float* packed, void* stream);
int mgb_in_edge_pack_norm(const float* gamma, const float* beta, float* packed, void* stream);
size_t mgb_in_edge_fwd_workspace(int64_t n_edges);
int mgb_in_edge_fwd(const float* e_features, float e_scale, const int32_t* perm, const float* pq, const int32_t* rowptr,
                    const int32_t* dst, const int32_t* src, int64_t n_nodes, int64_t n_edges, const float* packed, int precision,
                    float* agg, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the same edge function (autograd through models/magnet_gnn.py:70-90), recompute-based: the forward pass saves
 * nothing of size [E,128].  Two persistent tcgen05 passes over the edges (upper layers / lower layers; the weight gradients
 * accumulate in tensor memory, two per pass), see csrc/in_edge_bwd_tc.cu.  Inputs as mgb_in_edge_fwd plus dagg [N,128].
 *   dpq [N,256]: columns 0..127 = dP (summed by destination, fixed order); columns 128..255 zeroed (dQ = sum of dz0 rows by
 *                source: mgb_segment_sum_rows over the transposed plan, divided by e_scale);
 *   dz0 [E,128]: e_scale * d loss / d z0 in COO order (rows of the first Linear's pre-activation): d e_features = dz0 We,
 *                dWe = dz0^T e_features through mgb_linear_tc_bwd;
 *   dW [4][128][128], db [4][128]: gradients of edge_fn's Linears 1..4; dgamma, dbeta [128]: LayerNorm affine. */
size_t mgb_in_edge_bwd_workspace(int64_t n_edges);
int mgb_in_edge_bwd(const float* dagg, const float* e_features, float e_scale, const int32_t* perm, const float* pq,
                    const int32_t* rowptr, const int32_t* dst, const int32_t* src, int64_t n_nodes, int64_t n_edges,
                    const float* packed, int precision, float* dpq, float* dz0, float* dW, float* db, float* dgamma, float* dbeta,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * INR decoder.  Replaces the body of MAgNetGNN.continuous_decoder (models/magnet_gnn.py:254-280) given the
 * neighbour table of mgb_knn (:247).  a [B*L,128] = lr_encoded proj_head.weight[:, :128]^T + proj_head.bias
 * (the latent part of proj_head, factorised per low-resolution node); xlr [B,T,L]; lr_coords [B*L,d];
 * hr_coords [Q,d]; t [B,ldt] (first T columns used); wsmall = proj_head.weight + 128 with row stride ldw
 * (columns: input value, relative coordinates, time); idx [Q,k] global low-res row ids, ascending distance —
 * only neighbours 0 and 1 enter the blend.  mode: 0 'area', 1 'knn', 2 'sph'.
 *   fwd: z [Q,T,128].
 *   bwd: g [2Q,128] and sx [2Q,T] (per-neighbour contributions to d a / d xlr, to be summed per low-res node with
 *        mgb_segment_sum_rows), dwsmall [128, ldw-strided, d+2 columns] (+)= .
 * ------------------------------------------------------------------------------------------- */
int mgb_inr_decode_fwd(const float* a, const float* xlr, const float* lr_coords, const float* hr_coords, const float* t, int ldt,
                       const float* wsmall, int ldw, const int64_t* idx, int k, int64_t n_query, int nq_per_sample, int L, int T,
                       int d, int mode, float* z, void* stream);
size_t mgb_inr_decode_bwd_workspace(int64_t n_query);
int mgb_inr_decode_bwd(const float* a, const float* xlr, const float* lr_coords, const float* hr_coords, const float* t, int ldt,
                       const float* wsmall, int ldw, const int64_t* idx, int k, int64_t n_query, int nq_per_sample, int L, int T,
                       int d, int mode, const float* dz, float* g, float* sx, float* dwsmall, int accumulate, void* workspace,
                       size_t workspace_bytes, void* stream);

/* The whole query path of MAgNetGNN.forward (models/magnet_gnn.py:338-339: continuous_decoder + projector) in ONE launch,
 * forward only: per query the two nearest low-res nodes are searched in the grid hash of the low-res mesh (torch_cluster.knn
 * order, :247; or taken from idx [Q,k] when it is not NULL), their proj_head outputs are blended exactly as
 * mgb_inr_decode_fwd does, and every (query, time step) row goes straight through the projector MLP (packed as for
 * mgb_mlp_chain_fwd: n_layers Linears of width 128, ReLU between, n_out outputs) as tcgen05 tiles.  z [Q,T,128] never
 * exists.  y [Q*T, n_out].  ptr_x [n_samples+1]: offsets of the samples' low-res nodes (needed when idx is NULL; the grid
 * is built in the workspace by a few small launches).  grid_ready != 0: the workspace still holds the grid an earlier call built
 * over the same lr_coords (a rollout decodes every step over the same low-res mesh): nothing is rebuilt. */
size_t mgb_inr_decode_fused_workspace(int64_t n_lowres, int n_samples);
int mgb_inr_decode_fused(const float* a, const float* xlr, const float* lr_coords, const float* hr_coords, const float* t, int ldt,
                         const float* wsmall, int ldw, const int64_t* idx, int k, const int64_t* ptr_x, int n_samples, int64_t n_query,
                         int nq_per_sample, int L, int T, int d, int mode, int n_layers, const float* packed, int n_out, float* y,
                         int grid_ready, void* workspace, size_t workspace_bytes, void* stream);

/* InstanceNorm alone (PyG InstanceNorm, models/mpnn_2d.py:63,70) — exposed for tests. */
size_t mgb_instance_norm_workspace(int n_graphs, int max_nodes_per_graph);
int mgb_instance_norm_fwd(const float* x, const int64_t* gptr, int n_graphs, int max_nodes_per_graph, float* y,
                          float* rstd, void* workspace, size_t workspace_bytes, void* stream);

/* Test hook: one 128x128x128 bf16 tcgen05.mma tile, d[m][n] = sum_k A(m,k) B(n,k), through K-major
 * (x_mn_major = 0: a[m][k] / b[n][k]) or MN-major (x_mn_major = 1: a[k][m] / b[k][n]) shared-memory
 * descriptors over the same 128-byte-swizzled tile image.  lbo_mn / sbo_mn = 0 use the library defaults. */
int mgb_umma_selftest(const float* a, const float* b, int a_mn_major, int b_mn_major, int lbo_mn, int sbo_mn, float* d,
                      void* stream);

/* Test hooks for the integer primitives (stable radix sort, exclusive scan). */
size_t mgb_sort_workspace(int64_t n);
int mgb_sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                       int64_t n, int bits, void* workspace, size_t workspace_bytes, void* stream);
size_t mgb_scan_workspace(int64_t n);
int mgb_exclusive_scan_i32(const int32_t* in, int32_t* out /* n+1 */, int64_t n, void* workspace,
                           size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGNET_B200_H */
