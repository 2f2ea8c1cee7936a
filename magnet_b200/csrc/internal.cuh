// magnet_b200 — declarations shared between the kernel translation units and capi.cu
#pragma once
#include "dense.cuh"

namespace mgb {

const char* last_error();
void prof_enable(bool on);
int prof_collect(int id, double* total_ms, long long* count);

// graph.cu
size_t radius_workspace_bytes(int64_t n, int n_samples);
int radius_search(const float* pos, int64_t n, int d, const int64_t* ptr, int n_samples, double r, int max_num_neighbors,
                  int loop, int32_t* nbr, int32_t* deg, int32_t* rowptr, void* ws, size_t ws_bytes, cudaStream_t s);
int radius_emit(const int32_t* nbr, const int32_t* rowptr, int64_t n, int cap, int centre_row, int64_t n_edges,
                int64_t* edge_index, int32_t* col, cudaStream_t s);
size_t knn_workspace_bytes(int64_t nx, int n_samples);
int knn_search(const float* x, int64_t nx, const float* y, int64_t ny, int d, const int64_t* ptr_x, const int64_t* ptr_y,
               int n_samples, int k, int64_t* out_idx, float* out_dist, void* ws, size_t ws_bytes, cudaStream_t s);
size_t csr_plan_workspace_bytes(int64_t n_edges);
int csr_plan(const int64_t* agg, const int64_t* other, int64_t n_edges, int64_t n_nodes, int32_t* rowptr, int32_t* perm,
             int32_t* dst, int32_t* src, int32_t* rowptr_t, int32_t* pos_t, int* bad_flag, void* ws, size_t ws_bytes,
             cudaStream_t s);

// gnn_layer.cu
struct GnnLayerShape {
    int64_t n_nodes, n_edges;
    int tw, dp, nv;
    int n_graphs, max_nodes_per_graph;
    int precision;   // 0 fp32 FFMA, 1 tcgen05 bf16 hi/lo split (fp32 contract), 2 tcgen05 bf16
    int Kc() const { return 128 + tw + dp + nv; }
    int K1() const { return 256 + tw + dp + nv; }
    int K3() const { return 256 + nv; }
};
struct GnnFwdIO {
    const float *x, *u, *pos, *var;
    const int32_t *rowptr, *dstv, *srcv;
    const int64_t* gptr;
    const float* packed;
    const float *b2, *b3, *b4;
    float* y;
    float *pq, *agg, *y1_pre, *y2_pre, *rstd;
};
struct GnnBwdIO {
    const float* dy;
    const float *x, *u, *pos, *var;
    const float* y;
    const float *pq, *agg, *y1_pre, *y2_pre, *rstd;
    const int32_t *rowptr, *dstv, *srcv, *rowptr_t, *pos_t;
    const int64_t* gptr;
    const float* packed;
    const float *W2, *b2, *W3, *W4;
    float *dx, *du, *dpos, *dvar;
    float *dW1, *db1, *dW2, *db2, *dW3, *db3, *dW4, *db4;
    int accumulate_params;
};
size_t gnn_layer_packed_floats(int tw, int dp, int nv);
int gnn_layer_pack(const float* W1, const float* b1, const float* W2, const float* W3, const float* W4, int tw, int dp,
                   int nv, float* packed, cudaStream_t s);
size_t gnn_layer_fwd_workspace(int64_t n_nodes, int64_t n_edges, int n_graphs, int max_nodes);
int gnn_layer_fwd(const GnnLayerShape& sh, const GnnFwdIO& io, void* ws, size_t ws_bytes, cudaStream_t s);
size_t gnn_layer_bwd_workspace(int64_t n_nodes, int64_t n_edges, int tw, int dp, int nv, int n_graphs, int max_nodes);
int gnn_layer_bwd(const GnnLayerShape& sh, const GnnBwdIO& io, void* ws, size_t ws_bytes, cudaStream_t s);
size_t inorm_workspace_bytes(int n_graphs, int max_nodes);
int instance_norm_fwd(const float* x, const int64_t* gptr, int n_graphs, int max_nodes, float* y, float* rstd_out, void* ws,
                      size_t ws_bytes, cudaStream_t s);

// interaction.cu
int edge_combine_fwd(const float* P, const float* Q, const float* R, const int64_t* edge_index, int64_t n_edges, int act,
                     float* out, cudaStream_t s);
int relu_mask(const float* dout, const float* out, int64_t n, float* dz, cudaStream_t s);
int segment_sum_rows(const float* rows, int cols, const int32_t* rowptr, const int32_t* idx, int64_t n_nodes, int mean,
                     float* out, int ld_out, cudaStream_t s);
int gather_rows(const float* rows, const int64_t* index, const int32_t* rowptr, int64_t n_edges, float* out, cudaStream_t s);
int magnet_features(const float* u, int C, const float* x, int d, const float* t_last, int B, int64_t n_nodes,
                    const int64_t* edge_index, int64_t n_edges, float* nf, float* ef, cudaStream_t s);
struct InrArgs {
    const float* A;          // [B*L][128]
    const float* xlr;        // [B][T][L]
    const float* lr_coords;  // [B*L][d]
    const float* hr_coords;  // [Q][d]
    const float* t;          // [B][ldt]  (first T entries of each row are used)
    int ldt;
    const float* wsmall;     // proj_head.weight + C, row stride ldw
    int ldw;
    const int64_t* idx;      // [Q][k]  nearest low-res nodes (global row ids), ascending distance
    int k;
    int64_t n_query;
    int nq_per_sample, L, T, d, mode;    // mode 0 area, 1 knn, 2 sph
};
int inr_decode_fwd(const InrArgs& a, float* z, cudaStream_t s);
size_t inr_decode_bwd_workspace(int64_t n_query);
int inr_decode_bwd(const InrArgs& a, const float* dz, float* G, float* Sx, float* dwsmall, int ldw, int accumulate, void* ws,
                   size_t ws_bytes, cudaStream_t s);

// gnn_edge_tc.cu (tcgen05 edge kernels)
int pack_w2_image(const float* W2, void* img, cudaStream_t s);
size_t edge_fwd_tc_workspace(int64_t n_edges);
int launch_edge_fwd_tc(int precision, const float* pq, const int32_t* rowptr, const int32_t* dstv, const int32_t* srcv,
                       int64_t n_edges, const void* w2img, const float* b2, float* agg, void* ws, size_t ws_bytes,
                       cudaStream_t s);
size_t edge_bwd_tc_workspace(int64_t n_edges);
int launch_edge_bwd_tc(int precision, const float* pq, const int32_t* rowptr, const int32_t* dstv, const int32_t* srcv,
                       int64_t n_edges, const void* w2img, const float* b2, const float* dagg, int ld_dagg, float* dz1,
                       float* dpq, float* dW2, float* db2, int accumulate, void* ws, size_t ws_bytes, cudaStream_t s);
// linear_tc.cu (tcgen05 row-wise Linear, data and weight gradients)
struct LinTcArgs {
    const float* src[2]; int ld[2]; int nk;         // activation chunks of 128 columns
    const float* pre; int ldpre; int pre_act;       // chunk 0: x *= act'(pre)
    int self_act;                                   // chunk 0: x = act(x)
    const float* tsrc[3]; int tld[3]; int tk[3]; int kt;   // small-K tail (forward only), sum tk = kt <= 16
    const float* wtail; int wt_sn, wt_st;           // tail weight (n, t) at wtail[n*wt_sn + t*wt_st]
    const void* wimg;                               // weight tiles (hi|lo images), 64 KB each
    int nm, a_trans; int tile_of[2][2];             // tile index used by (output block m, chunk kc)
    const float* bias; int act;
    int* range_flag;                                // fp16 split only: set to 1 by a producer that meets |x| >= 32768 (host-mapped)
    int n_out;                                      // valid output columns (0 = all 128*nm): the rest is neither biased nor stored
    const float* residual; int ldr; int res_blocks;   // residual added to output block m if bit m of res_blocks is set (0 = all blocks)
    float* y; int ldy; float* y_pre; int ldyp;
    int64_t rows;
};
struct WgradTcArgs {
    const float* dy; int lddy; int ny;                    // Y'_y = dy[:, 128 y : 128 y + 128] * act'(y_pre[:, same])
    const float* y_pre; int ldyp; int y_act;
    int nx; const float* x[2]; int ldx[2]; int x_act[2];  // 128-column sources (optionally act(x))
    int tail;                                             // 1: a tail tile [small-K columns | 0...] follows the sources
    const float* tsrc[3]; int tld[3]; int tk[3]; int kt;  // small-K columns, sum tk = kt <= 16
    float* db[2]; int db_accumulate;                      // optional bias gradients: column sums of Y'_y (exact fp32 adds)
    int64_t rows;
    float* partial;                                       // filled by the launcher
    float* bias_partial;                                  // filled by the launcher: [grid][producer warps][ny][128]
};
struct WgradTcOut {          // destination of accumulator (y, x): dw[n*lddw + k], n < n_valid, k < k_valid
    float* dw; int lddw; int n_valid, k_valid; int accumulate;
};
// node_update_tc.cu (update_net_1 + update_net_2 + residual of GNN_Layer in one launch, forward)
struct NodeUpdateArgs {
    const float* x; const float* agg; int64_t rows;        // [N,128] each
    const float* var; int nv;                               // [N,nv], nv <= 4
    const void* wimg;                                       // W3[:, 0:128], W3[:, 128:256], W4 in tensor-memory order (pack_weight_tmem_bf16), 64 KB each
    const float* w3tail; int wt_sn, wt_st;                  // W3[n, 256 + t] at w3tail[n*wt_sn + t*wt_st]
    const float* b3; const float* b4;
    float* y1_pre; float* y2_pre; float* out;               // [N,128]
};
int pack_weight_tmem_bf16(const float* W, int ld, int c0, void* out, cudaStream_t s, int trans = 0);
int launch_node_update_tc(int precision, const NodeUpdateArgs& a, cudaStream_t s);
int pack_weight_tile(const float* W, int ld, int n_rows, int n_cols, int r0, int c0, void* img, cudaStream_t s, int f16 = 0);
int pack_weight_tmem(const float* W, int ld, int n_rows, int n_cols, void* out, cudaStream_t s);
int launch_linear_tc(int precision, const LinTcArgs& a, cudaStream_t s);
// linear_ts.cu: the same contract with the weights in tensor memory (a.wimg = tiles in pack_weight_tmem_bf16 order), for the
// shapes linear_ts_covers accepts (no activation, no pre-activation copy, all output columns valid, weights not transposed)
bool linear_ts_covers(const LinTcArgs& a);
int launch_linear_ts(int precision, const LinTcArgs& a, cudaStream_t s);
size_t wgrad_tc_workspace(int64_t rows);
int launch_wgrad_tc(int precision, WgradTcArgs a, const WgradTcOut* outs /*[ny][nx + tail]*/, void* ws, size_t ws_bytes,
                    cudaStream_t s);
int launch_tail_dgrad(const float* dy, int lddy, int ny, const float* pre, int ldpre, int act, const float* wtail, int wt_sn,
                      int wt_st, int kt, int64_t rows, float* out, int ldo, cudaStream_t s);
#ifdef MGB_TIMELINE
int set_timeline_buffer(long long* p);
int set_ie_timeline_buffer(long long* p);
int set_ib_timeline_buffer(long long* p);
int set_lt_timeline_buffer(long long* p);
int set_nu_timeline_buffer(long long* p);
#endif
// mlp_chain_tc.cu (a whole 128-wide MLP in one launch, forward only)
struct CellPoint;
struct GridParams;
struct InrFuseArgs {                 // fused INR decoder: the rows of the chain are (query, time step) pairs computed in place
    const CellPoint* pts; const int32_t* cell_start; const GridParams* gp;     // grid of the low-res nodes (grid.cuh), or
    const int64_t* idx; int k;       // a precomputed neighbour table [Q][k] (then no search)
    const float* A;                  // [B*L][128] latent part of proj_head per low-res node (+ bias)
    const float* xlr;                // [B][T][L]
    const float* lr_coords;          // [B*L][d]
    const float* hr_coords;          // [Q][d]
    const float* t; int ldt;         // [B][ldt], first T entries used
    const float* wsmall; int ldw;    // proj_head.weight + 128: columns input value, relative coordinates, time
    int64_t n_query; int nq_per_sample, L, T, d, mode;
};
struct MlpChainArgs {
    InrFuseArgs inr;                 // used by the fused decoder only
    const float* x; int ldx; int64_t rows;
    int n_layers;                    // Linear layers, all with 128 inputs; 128 outputs except the last (n_out <= 128)
    const void* wimg;                // [n_layers] fp16 hi | lo pairs of W_l [out, 128] in tensor-memory order (pack_weight_tmem), 64 KB each
    const float* bias;               // [n_layers][128], zero-padded
    int act;                         // activation after every layer but the last
    int in_act;                      // activation applied to x on load (0 none, 1 ReLU)
    int n_out; float* y; int ldy;
    int* range_flag;                 // raised when an operand leaves the fp16 range
};
int launch_mlp_chain_tc(const MlpChainArgs& a, cudaStream_t s);
int launch_inr_decode_fused(const MlpChainArgs& a, cudaStream_t s);      // a.inr filled; a.rows = n_query * T
// in_edge_tc.cu (fused InteractionNetwork edge function: gather + 5-layer MLP + LayerNorm + segmented mean)
size_t in_edge_fwd_workspace(int64_t n_edges);
int launch_in_edge_fwd(int precision, const float* e, float e_scale, const int32_t* perm, const float* pq, const int32_t* rowptr,
                       const int32_t* dstv, const int32_t* srcv, int64_t n_nodes, int64_t n_edges, const float* packed, float* agg,
                       int* range_flag, void* ws, size_t ws_bytes, cudaStream_t s);
// in_edge_bwd_tc.cu (its backward: two recompute passes, weight gradients resident in tensor memory)
size_t in_edge_bwd_workspace(int64_t n_edges);
int launch_in_edge_bwd(int precision, const float* dagg, const float* e, float e_scale, const int32_t* perm, const float* pq,
                       const int32_t* rowptr, const int32_t* dstv, const int32_t* srcv, int64_t n_nodes, int64_t n_edges,
                       const float* packed, float* dpq, float* dz0, float* dW, float* db, float* dgamma, float* dbeta, int* range_flag,
                       void* ws, size_t ws_bytes, cudaStream_t s);
// decoder.cu (temporal-bundling Conv1d decoder + Euler update of MP-PDE, one launch per direction)
struct DecArgs {
    const float* h;        // [N][128]
    const float* u; int ldu, u_col;     // u[:, u_col] = last input value
    const float* w1;       // [8][k1]
    const float* b1;       // [8]
    const float* w2;       // [8][k2]   (Conv1d(8,1,k2).weight [1,8,k2])
    const float* b2;       // [1]
    const float* dt;       // device scalar
    int k1, s1, k2, L1, tw, act;
    int64_t n;
};
int decoder_fwd(DecArgs a, int hidden, float* out, cudaStream_t s);
size_t decoder_bwd_workspace(int64_t n);
int decoder_bwd(DecArgs a, int hidden, const float* dout, float* dh, float* du, int lddu, float* dw1, float* db1, float* dw2, float* db2,
                void* ws, size_t ws_bytes, cudaStream_t s);
// small_linear.cu (Linears with <= 16 input features and 128 outputs: streaming kernels, no GEMM tiles)
bool small_linear_ok(int in_features, int out_features);
int small_linear_fwd(const float* x, int64_t rows, int K, const float* wt, const float* bias, int act, float* y, float* y_pre, cudaStream_t s);
size_t small_linear_bwd_workspace(int64_t rows, int K);
int small_linear_bwd(const float* dy, const float* y_pre, int act, const float* x, int64_t rows, int K, const float* w, float* dx, float* dw,
                     float* db, int accumulate, void* ws, size_t ws_bytes, cudaStream_t s);
// optim.cu
int adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
              double weight_decay, int64_t step, double grad_scale, cudaStream_t s);
// umma_selftest.cu
int umma_selftest(const float* a, const float* b, int a_mn, int b_mn, int lbo_mn, int sbo_mn, float* d, cudaStream_t s);

}  // namespace mgb
