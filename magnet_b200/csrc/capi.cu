// magnet_b200 — extern "C" entry points (see include/magnet_b200.h for the contract).
#include <mutex>
#include "internal.cuh"
#include "grid.cuh"
#include "../../include/magnet_b200.h"

using namespace mgb;

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int mgb_abi_version(void) { return MGB_ABI_VERSION; }
const char* mgb_last_error(void) { return mgb::last_error(); }
int64_t mgb_launch_count(void) { return (int64_t)mgb::launch_count(); }

int mgb_profile_enable(int on) { mgb::prof_enable(on != 0); return MGB_OK; }
int mgb_profile_collect(int kernel_id, double* total_ms, int64_t* count) {
    MGB_REQUIRE(kernel_id >= 0 && kernel_id < mgb::PROF_COUNT, "profile_collect: unknown kernel id %d", kernel_id);
    long long c = 0;
    int rc = mgb::prof_collect(kernel_id, total_ms, &c);
    *count = c;
    return rc;
}

size_t mgb_radius_graph_workspace(int64_t n, int n_samples) { return radius_workspace_bytes(n, n_samples); }

int mgb_radius_graph_search(const float* pos, int64_t n, int d, const int64_t* ptr, int n_samples, double r,
                            int max_num_neighbors, int loop, int32_t* nbr, int32_t* deg, int32_t* rowptr,
                            void* workspace, size_t workspace_bytes, void* stream) {
    return radius_search(pos, n, d, ptr, n_samples, r, max_num_neighbors, loop, nbr, deg, rowptr, workspace,
                         workspace_bytes, STREAM(stream));
}

int mgb_radius_graph_emit(const int32_t* nbr, const int32_t* rowptr, int64_t n, int cap, int centre_row,
                          int64_t n_edges, int64_t* edge_index, int32_t* col, void* stream) {
    return radius_emit(nbr, rowptr, n, cap, centre_row, n_edges, edge_index, col, STREAM(stream));
}

size_t mgb_knn_workspace(int64_t nx, int n_samples) { return knn_workspace_bytes(nx, n_samples); }

int mgb_knn(const float* x, int64_t nx, const float* y, int64_t ny, int d, const int64_t* ptr_x, const int64_t* ptr_y,
            int n_samples, int k, int64_t* out_idx, float* out_dist, void* workspace, size_t workspace_bytes,
            void* stream) {
    return knn_search(x, nx, y, ny, d, ptr_x, ptr_y, n_samples, k, out_idx, out_dist, workspace, workspace_bytes,
                      STREAM(stream));
}

size_t mgb_csr_plan_workspace(int64_t n_edges) { return csr_plan_workspace_bytes(n_edges); }

int mgb_csr_plan(const int64_t* agg, const int64_t* other, int64_t n_edges, int64_t n_nodes, int32_t* rowptr,
                 int32_t* perm, int32_t* dst, int32_t* src, int32_t* rowptr_t, int32_t* pos_t, int32_t* bad_flag,
                 void* workspace, size_t workspace_bytes, void* stream) {
    return csr_plan(agg, other, n_edges, n_nodes, rowptr, perm, dst, src, rowptr_t, pos_t, bad_flag, workspace,
                    workspace_bytes, STREAM(stream));
}

size_t mgb_gnn_layer_packed_floats(int tw, int dp, int nv) { return gnn_layer_packed_floats(tw, dp, nv); }

int mgb_gnn_layer_pack(const float* W1, const float* b1, const float* W2, const float* W3, const float* W4, int tw,
                       int dp, int nv, float* packed, void* stream) {
    MGB_REQUIRE(tw >= 1 && dp >= 1 && nv >= 1, "gnn_layer_pack: tw, dp, nv must be positive");
    return gnn_layer_pack(W1, b1, W2, W3, W4, tw, dp, nv, packed, STREAM(stream));
}

size_t mgb_gnn_layer_fwd_workspace(int64_t n_nodes, int64_t n_edges, int n_graphs, int max_nodes_per_graph) {
    return gnn_layer_fwd_workspace(n_nodes, n_edges, n_graphs, max_nodes_per_graph);
}

int mgb_gnn_layer_fwd(int64_t n_nodes, int64_t n_edges, int tw, int dp, int nv, int n_graphs, int max_nodes_per_graph,
                      const float* x, const float* u, const float* pos, const float* var, const int32_t* rowptr,
                      const int32_t* dst, const int32_t* src, const int64_t* gptr, const float* packed,
                      const float* b2, const float* b3, const float* b4, float* y, float* pq, float* agg,
                      float* y1_pre, float* y2_pre, float* rstd, int precision, void* workspace, size_t workspace_bytes,
                      void* stream) {
    GnnLayerShape sh{n_nodes, n_edges, tw, dp, nv, n_graphs, max_nodes_per_graph, precision};
    GnnFwdIO io{x, u, pos, var, rowptr, dst, src, gptr, packed, b2, b3, b4, y, pq, agg, y1_pre, y2_pre, rstd};
    return gnn_layer_fwd(sh, io, workspace, workspace_bytes, STREAM(stream));
}

// packed block of the stand-alone node update: three tensor-memory weight images (H*H floats each) + the var columns of W3 [128][4]
size_t mgb_gnn_node_update_packed_floats(void) { return (size_t)3 * 128 * 128 + 128 * 4; }
int mgb_gnn_node_update_pack(const float* W3, const float* W4, int nv, float* packed, void* stream) {
    MGB_REQUIRE(nv >= 0 && nv <= 4, "node_update: at most 4 var columns");
    cudaStream_t s = STREAM(stream);
    MGB_TRY(pack_weight_tmem_bf16(W3, 256 + nv, 0, packed, s));
    MGB_TRY(pack_weight_tmem_bf16(W3, 256 + nv, 128, packed + 128 * 128, s));
    MGB_TRY(pack_weight_tmem_bf16(W4, 128, 0, packed + 2 * 128 * 128, s));
    MGB_CUDA(cudaMemsetAsync(packed + 3 * 128 * 128, 0, 128 * 4 * sizeof(float), s));
    if (nv > 0)
        MGB_CUDA(cudaMemcpy2DAsync(packed + 3 * 128 * 128, 4 * sizeof(float), W3 + 256, (size_t)(256 + nv) * sizeof(float), (size_t)nv * sizeof(float),
                                   128, cudaMemcpyDeviceToDevice, s));
    return MGB_OK;
}
int mgb_gnn_node_update_fwd(const float* x, const float* agg, const float* var, int nv, int64_t n_nodes, const float* packed,
                            const float* b3, const float* b4, float* y1_pre, float* y2_pre, float* out, int precision,
                            void* stream) {
    MGB_REQUIRE(precision == 1 || precision == 2, "node_update: precision must be 1 (bf16 hi/lo split) or 2 (bf16)");
    NodeUpdateArgs a{};
    a.x = x; a.agg = agg; a.rows = n_nodes; a.var = var; a.nv = nv;
    a.wimg = packed; a.w3tail = packed + 3 * 128 * 128; a.wt_sn = 4; a.wt_st = 1;
    a.b3 = b3; a.b4 = b4; a.y1_pre = y1_pre; a.y2_pre = y2_pre; a.out = out;
    return launch_node_update_tc(precision, a, STREAM(stream));
}

size_t mgb_gnn_layer_bwd_workspace(int64_t n_nodes, int64_t n_edges, int tw, int dp, int nv, int n_graphs,
                                   int max_nodes_per_graph) {
    return gnn_layer_bwd_workspace(n_nodes, n_edges, tw, dp, nv, n_graphs, max_nodes_per_graph);
}

int mgb_gnn_layer_bwd(int64_t n_nodes, int64_t n_edges, int tw, int dp, int nv, int n_graphs, int max_nodes_per_graph,
                      const float* dy, const float* x, const float* u, const float* pos, const float* var,
                      const float* y, const float* pq, const float* agg, const float* y1_pre, const float* y2_pre,
                      const float* rstd, const int32_t* rowptr, const int32_t* dst, const int32_t* src,
                      const int32_t* rowptr_t, const int32_t* pos_t, const int64_t* gptr, const float* packed,
                      const float* W2, const float* b2, const float* W3, const float* W4, float* dx, float* du,
                      float* dpos, float* dvar, float* dW1, float* db1, float* dW2, float* db2, float* dW3,
                      float* db3, float* dW4, float* db4, int accumulate_params, int precision, void* workspace,
                      size_t workspace_bytes, void* stream) {
    GnnLayerShape sh{n_nodes, n_edges, tw, dp, nv, n_graphs, max_nodes_per_graph, precision};
    GnnBwdIO io{dy, x, u, pos, var, y, pq, agg, y1_pre, y2_pre, rstd, rowptr, dst, src, rowptr_t, pos_t, gptr, packed,
                W2, b2, W3, W4, dx, du, dpos, dvar, dW1, db1, dW2, db2, dW3, db3, dW4, db4, accumulate_params};
    return gnn_layer_bwd(sh, io, workspace, workspace_bytes, STREAM(stream));
}

int mgb_transpose(const float* in, int rows, int cols, float* out, void* stream) {
    return launch_transpose(in, rows, cols, cols, out, rows, STREAM(stream));
}

int mgb_linear_fwd(const float* x, int64_t rows, int in_features, int out_features, const float* wt, const float* bias,
                   int act, const float* residual, float* y, float* y_pre, void* stream) {
    MGB_REQUIRE(rows >= 0 && rows < ((int64_t)1 << 31), "linear_fwd: row count out of range");
    MGB_REQUIRE(act >= 0 && act <= 2, "linear_fwd: unknown activation %d", act);
    if (small_linear_ok(in_features, out_features) && residual == nullptr)      // a handful of inputs: one streaming kernel, no GEMM tiles
        return small_linear_fwd(x, rows, in_features, wt, bias, act, y, y_pre, STREAM(stream));
    GemmArgs g{};
    g.a.p[0] = x; g.a.ld[0] = in_features; g.a.k[0] = in_features; g.a.nseg = 1;
    g.b = wt; g.ldb = out_features; g.bias = bias;
    g.residual = residual; g.ldr = out_features;
    g.c = y; g.ldc = out_features; g.c_pre = y_pre; g.ldcp = out_features; g.act = act;
    g.M = (int)rows; g.N = out_features; g.K = in_features;
    return launch_gemm(g, STREAM(stream));
}

// ---- nn.Linear on the tensor cores (linear_tc.cu): in_features 128 or 256, out_features <= 256, at most two weight tiles
static bool linear_tc_shape(int in_features, int out_features, int* nk, int* nm) {
    if (in_features != 128 && in_features != 256) return false;
    if (out_features < 1 || out_features > 256) return false;
    *nk = in_features / 128;
    *nm = (out_features + 127) / 128;
    return *nk * *nm <= 2;
}

static bool linear_ts_bwd_shape(int in_features, int out_features) { return out_features == 128 && (in_features == 128 || in_features == 256); }

size_t mgb_linear_tc_packed_floats(int in_features, int out_features) {
    int nk, nm;
    if (!linear_tc_shape(in_features, out_features, &nk, &nm)) return 0;
    // per tile: two 16-bit images (hi | lo) of 128 x 128; then, for the shapes mgb_linear_tc_bwd covers, the blocks of W^T in
    // tensor-memory order (bf16 hi | lo, linear_ts.cu) for the data gradient
    return (size_t)nk * nm * 2 * 128 * 128 / 2 + (linear_ts_bwd_shape(in_features, out_features) ? (size_t)nk * 128 * 128 : 0);
}

int mgb_linear_tc_pack(const float* W, int ldw, int in_features, int out_features, int precision, float* packed, void* stream) {
    int nk, nm;
    MGB_REQUIRE(linear_tc_shape(in_features, out_features, &nk, &nm), "linear_tc_pack: unsupported shape %d -> %d", in_features, out_features);
    for (int m = 0; m < nm; ++m)
        for (int kc = 0; kc < nk; ++kc)
            MGB_TRY(pack_weight_tile(W, ldw, out_features, in_features, m * 128, kc * 128, packed + (size_t)(m * nk + kc) * 128 * 128, STREAM(stream), precision == 3));
    if (linear_ts_bwd_shape(in_features, out_features) && precision != 3)
        for (int kc = 0; kc < nk; ++kc)      // A_kc[k][n] = W[n][kc*128 + k]
            MGB_TRY(pack_weight_tmem_bf16(W, ldw, kc * 128, packed + (size_t)(nk * nm + kc) * 128 * 128, STREAM(stream), 1));
    return MGB_OK;
}

// fp16 range guard of the fp16-split Linear: a host-mapped flag the producers raise when they meet |x| >= 32768.  It is
// checked (without synchronising) at the start of every later call, so an out-of-range input surfaces as an error on one
// of the next calls instead of silently turning into infinities.
static int* f16_range_flag(int** dev_ptr) {
    static int* host = nullptr;
    static int* dev = nullptr;
    static std::once_flag once;          // forward runs on the caller's thread, backward on PyTorch's autograd thread
    std::call_once(once, [] {
        if (cudaHostAlloc((void**)&host, sizeof(int), cudaHostAllocMapped) != cudaSuccess) { host = nullptr; return; }
        *host = 0;
        if (cudaHostGetDevicePointer((void**)&dev, host, 0) != cudaSuccess) dev = nullptr;
    });
    if (dev_ptr) *dev_ptr = dev;
    return host;
}

// 1 (and the flag is cleared) when a fp16-split kernel has met |x| >= 32768 since the last check.  Reads host memory only; for an
// answer that covers the calls issued so far the caller synchronises their stream first (a natural point: before results are used).
int mgb_f16_range_check(void) {
    volatile int* host_flag = f16_range_flag(nullptr);
    if (host_flag && *host_flag) {
        *host_flag = 0;
        return 1;
    }
    return 0;
}

int mgb_linear_tc_fwd(const float* x, int64_t rows, int in_features, int out_features, const float* packed, const float* bias,
                      int act, const float* residual, float* y, float* y_pre, int precision, void* stream) {
    int nk, nm;
    MGB_REQUIRE(linear_tc_shape(in_features, out_features, &nk, &nm), "linear_tc_fwd: unsupported shape %d -> %d", in_features, out_features);
    MGB_REQUIRE(act >= 0 && act <= 2, "linear_tc_fwd: unknown activation %d", act);
    MGB_REQUIRE(precision >= 1 && precision <= 3, "linear_tc_fwd: precision must be 1 (bf16 hi/lo split), 2 (bf16) or 3 (fp16 hi/lo split)");
    LinTcArgs a{};
    if (precision == 3) {
        int* dev_flag = nullptr;
        volatile int* host_flag = f16_range_flag(&dev_flag);
        if (host_flag && *host_flag) {
            *host_flag = 0;
            MGB_REQUIRE(false, "linear_tc_fwd: an earlier fp16-split Linear met |x| >= 32768 (fp16 range); use the fp32 path (set_linear_tc(False)) for this data");
        }
        a.range_flag = dev_flag;
    }
    for (int kc = 0; kc < nk; ++kc) { a.src[kc] = x + kc * 128; a.ld[kc] = in_features; }
    a.nk = nk; a.nm = nm;
    for (int m = 0; m < nm; ++m)
        for (int kc = 0; kc < nk; ++kc) a.tile_of[m][kc] = m * nk + kc;
    a.wimg = packed; a.bias = bias; a.act = act; a.n_out = out_features;
    a.residual = residual; a.ldr = out_features;
    a.y = y; a.ldy = out_features; a.y_pre = y_pre; a.ldyp = out_features; a.rows = rows;
    return launch_linear_tc(precision, a, STREAM(stream));
}

// the same Linear over cat([x0, x1], dim = 1) of two 128-column tensors without materialising the concatenation
int mgb_linear_tc_fwd2(const float* x0, int ld0, const float* x1, int ld1, int64_t rows, int out_features, const float* packed,
                       const float* bias, int act, float* y, int precision, void* stream) {
    int nk, nm;
    MGB_REQUIRE(linear_tc_shape(256, out_features, &nk, &nm), "linear_tc_fwd2: unsupported shape 256 -> %d", out_features);
    MGB_REQUIRE(act >= 0 && act <= 2, "linear_tc_fwd2: unknown activation %d", act);
    MGB_REQUIRE(precision >= 1 && precision <= 3, "linear_tc_fwd2: precision must be 1, 2 or 3");
    MGB_REQUIRE(ld0 >= 128 && ld1 >= 128 && ld0 % 4 == 0 && ld1 % 4 == 0, "linear_tc_fwd2: row strides must be multiples of 4 floats, >= 128");
    LinTcArgs a{};
    if (precision == 3) {
        int* dev_flag = nullptr;
        volatile int* host_flag = f16_range_flag(&dev_flag);
        if (host_flag && *host_flag) {
            *host_flag = 0;
            MGB_REQUIRE(false, "linear_tc_fwd2: an earlier fp16-split Linear met |x| >= 32768 (fp16 range); use the fp32 path (set_linear_tc(False)) for this data");
        }
        a.range_flag = dev_flag;
    }
    a.src[0] = x0; a.ld[0] = ld0; a.src[1] = x1; a.ld[1] = ld1;
    a.nk = nk; a.nm = nm;
    for (int m = 0; m < nm; ++m)
        for (int kc = 0; kc < nk; ++kc) a.tile_of[m][kc] = m * nk + kc;
    a.wimg = packed; a.bias = bias; a.act = act; a.n_out = out_features;
    a.y = y; a.ldy = out_features; a.rows = rows;
    return launch_linear_tc(precision, a, STREAM(stream));
}

// ---- a whole 128-wide MLP in one launch (mlp_chain_tc.cu): packed = [n_layers][hi | lo images] | [n_layers][128] biases
size_t mgb_mlp_chain_packed_floats(int n_layers) {
    if (n_layers < 1 || n_layers > 8) return 0;
    return (size_t)n_layers * 128 * 128 + (size_t)n_layers * 128;
}

int mgb_mlp_chain_pack_layer(const float* W, int ldw, int out_features, const float* bias, int layer, int n_layers, float* packed,
                             void* stream) {
    MGB_REQUIRE(n_layers >= 1 && n_layers <= 8 && layer >= 0 && layer < n_layers, "mlp_chain_pack_layer: layer %d of %d", layer, n_layers);
    MGB_REQUIRE(out_features >= 1 && out_features <= 128, "mlp_chain_pack_layer: 1 <= out_features <= 128 (got %d)", out_features);
    MGB_TRY(pack_weight_tmem(W, ldw, out_features, 128, packed + (size_t)layer * 128 * 128, STREAM(stream)));
    float* b = packed + (size_t)n_layers * 128 * 128 + (size_t)layer * 128;
    MGB_CUDA(cudaMemsetAsync(b, 0, 128 * sizeof(float), STREAM(stream)));
    if (bias) MGB_CUDA(cudaMemcpyAsync(b, bias, (size_t)out_features * sizeof(float), cudaMemcpyDeviceToDevice, STREAM(stream)));
    return MGB_OK;
}

int mgb_mlp_chain_fwd(const float* x, int ldx, int64_t rows, int n_layers, const float* packed, int act, int in_act, int n_out, float* y,
                      int ldy, void* stream) {
    MGB_REQUIRE(act >= 0 && act <= 2 && (in_act == 0 || in_act == 1), "mlp_chain_fwd: unknown activation");
    MGB_REQUIRE(n_layers >= 1 && n_layers <= 8, "mlp_chain_fwd: 1..8 layers (got %d)", n_layers);
    MlpChainArgs a{};
    int* dev_flag = nullptr;
    volatile int* host_flag = f16_range_flag(&dev_flag);
    if (host_flag && *host_flag) {
        *host_flag = 0;
        MGB_REQUIRE(false, "mlp_chain_fwd: an earlier fp16-split Linear met |x| >= 32768 (fp16 range); use the fp32 path (set_linear_tc(False)) for this data");
    }
    a.range_flag = dev_flag;
    a.x = x; a.ldx = ldx; a.rows = rows; a.n_layers = n_layers;
    a.wimg = packed; a.bias = packed + (size_t)n_layers * 128 * 128;
    a.act = act; a.in_act = in_act; a.n_out = n_out; a.y = y; a.ldy = ldy;
    return launch_mlp_chain_tc(a, STREAM(stream));
}

// ---- fused InteractionNetwork edge function (in_edge_tc.cu): packed = [5][hi | lo images] | [5][128] biases | gamma | beta
size_t mgb_in_edge_packed_floats(void) { return (size_t)5 * 128 * 128 + 5 * 128 + 256; }

int mgb_in_edge_pack_layer(const float* W, int ldw, const float* bias, int layer, int precision, float* packed, void* stream) {
    MGB_REQUIRE(layer >= 0 && layer < 5, "in_edge_pack_layer: layer %d of 5", layer);
    MGB_TRY(pack_weight_tile(W, ldw, 128, 128, 0, 0, packed + (size_t)layer * 128 * 128, STREAM(stream), precision == 2 ? 0 : 1));
    float* b = packed + (size_t)5 * 128 * 128 + (size_t)layer * 128;
    if (bias) MGB_CUDA(cudaMemcpyAsync(b, bias, 128 * sizeof(float), cudaMemcpyDeviceToDevice, STREAM(stream)));
    else MGB_CUDA(cudaMemsetAsync(b, 0, 128 * sizeof(float), STREAM(stream)));
    return MGB_OK;
}

int mgb_in_edge_pack_norm(const float* gamma, const float* beta, float* packed, void* stream) {
    float* g = packed + (size_t)5 * 128 * 128 + 5 * 128;
    MGB_CUDA(cudaMemcpyAsync(g, gamma, 128 * sizeof(float), cudaMemcpyDeviceToDevice, STREAM(stream)));
    MGB_CUDA(cudaMemcpyAsync(g + 128, beta, 128 * sizeof(float), cudaMemcpyDeviceToDevice, STREAM(stream)));
    return MGB_OK;
}

size_t mgb_in_edge_fwd_workspace(int64_t n_edges) { return in_edge_fwd_workspace(n_edges); }

int mgb_in_edge_fwd(const float* e_features, float e_scale, const int32_t* perm, const float* pq, const int32_t* rowptr,
                    const int32_t* dst, const int32_t* src, int64_t n_nodes, int64_t n_edges, const float* packed, int precision,
                    float* agg, void* workspace, size_t workspace_bytes, void* stream) {
    int* dev_flag = nullptr;
    if (precision != 2) {
        volatile int* host_flag = f16_range_flag(&dev_flag);
        if (host_flag && *host_flag) {
            *host_flag = 0;
            MGB_REQUIRE(false, "in_edge_fwd: an earlier fp16-split kernel met |x| >= 32768 (fp16 range); use the fp32 path (set_linear_tc(False)) for this data");
        }
    }
    return launch_in_edge_fwd(precision, e_features, e_scale, perm, pq, rowptr, dst, src, n_nodes, n_edges, packed, agg, dev_flag,
                              workspace, workspace_bytes, STREAM(stream));
}

int mgb_magnet_features(const float* u, int n_chan, const float* x, int d, const float* t_last, int n_samples, int64_t n_nodes,
                        const int64_t* edge_index, int64_t n_edges, float* node_features, float* edge_features, void* stream) {
    return magnet_features(u, n_chan, x, d, t_last, n_samples, n_nodes, edge_index, n_edges, node_features, edge_features, STREAM(stream));
}

size_t mgb_in_edge_bwd_workspace(int64_t n_edges) { return in_edge_bwd_workspace(n_edges); }

int mgb_in_edge_bwd(const float* dagg, const float* e_features, float e_scale, const int32_t* perm, const float* pq,
                    const int32_t* rowptr, const int32_t* dst, const int32_t* src, int64_t n_nodes, int64_t n_edges,
                    const float* packed, int precision, float* dpq, float* dz0, float* dW, float* db, float* dgamma, float* dbeta,
                    void* workspace, size_t workspace_bytes, void* stream) {
    int* dev_flag = nullptr;
    if (precision != 2) {
        volatile int* host_flag = f16_range_flag(&dev_flag);
        if (host_flag && *host_flag) {
            *host_flag = 0;
            MGB_REQUIRE(false, "in_edge_bwd: an earlier fp16-split kernel met |x| >= 32768 (fp16 range); use the fp32 path (set_linear_tc(False)) for this data");
        }
    }
    return launch_in_edge_bwd(precision, dagg, e_features, e_scale, perm, pq, rowptr, dst, src, n_nodes, n_edges, packed, dpq, dz0, dW,
                              db, dgamma, dbeta, dev_flag, workspace, workspace_bytes, STREAM(stream));
}

// ---- fused INR decoder (mlp_chain_tc.cu, MODE 1): search + gather + proj_head + blend + projector in one launch
size_t mgb_inr_decode_fused_workspace(int64_t n_lowres, int n_samples) { return grid_workspace_bytes(n_lowres, n_samples) + 1024; }

int mgb_inr_decode_fused(const float* a, const float* xlr, const float* lr_coords, const float* hr_coords, const float* t, int ldt,
                         const float* wsmall, int ldw, const int64_t* idx, int k, const int64_t* ptr_x, int n_samples, int64_t n_query,
                         int nq_per_sample, int L, int T, int d, int mode, int n_layers, const float* packed, int n_out, float* y,
                         int grid_ready, void* workspace, size_t workspace_bytes, void* stream) {
    MGB_REQUIRE(n_layers >= 1 && n_layers <= 8, "inr_decode_fused: 1..8 projector layers (got %d)", n_layers);
    MGB_REQUIRE(d == 1 || d == 2, "inr_decode_fused: coordinate dimension must be 1 or 2");
    MlpChainArgs c{};
    int* dev_flag = nullptr;
    volatile int* host_flag = f16_range_flag(&dev_flag);
    if (host_flag && *host_flag) {
        *host_flag = 0;
        MGB_REQUIRE(false, "inr_decode_fused: an earlier fp16-split kernel met |x| >= 32768 (fp16 range); use the fp32 path (set_linear_tc(False)) for this data");
    }
    c.range_flag = dev_flag;
    InrFuseArgs& f = c.inr;
    if (idx == nullptr) {
        MGB_REQUIRE(ptr_x != nullptr, "inr_decode_fused: the in-kernel search needs the sample offsets of the low-res nodes");
        GridParams* gp; CellPoint* pts; int32_t* cell_start;
        MGB_TRY(build_grid_ws(lr_coords, (int64_t)n_samples * L, d, ptr_x, n_samples, 0.f, 2.0f, workspace, workspace_bytes, &gp, &pts,
                              &cell_start, STREAM(stream), grid_ready));
        f.gp = gp; f.pts = pts; f.cell_start = cell_start;
    }
    f.idx = idx; f.k = k; f.A = a; f.xlr = xlr; f.lr_coords = lr_coords; f.hr_coords = hr_coords; f.t = t; f.ldt = ldt;
    f.wsmall = wsmall; f.ldw = ldw; f.n_query = n_query; f.nq_per_sample = nq_per_sample; f.L = L; f.T = T; f.d = d; f.mode = mode;
    c.rows = n_query * T; c.n_layers = n_layers; c.wimg = packed; c.bias = packed + (size_t)n_layers * 128 * 128;
    c.act = ACT_RELU; c.in_act = ACT_NONE; c.n_out = n_out; c.y = y; c.ldy = n_out;
    return launch_inr_decode_fused(c, STREAM(stream));
}

// ---- temporal-bundling decoder + Euler update (decoder.cu)
static DecArgs dec_args(const float* h, int64_t n, const float* u, int ldu, int u_col, const float* w1, const float* b1, int k1, int stride1,
                        const float* w2, const float* b2, int k2, int time_window, int act, const float* dt) {
    DecArgs a{};
    a.h = h; a.u = u; a.ldu = ldu; a.u_col = u_col; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.dt = dt;
    a.k1 = k1; a.s1 = stride1; a.k2 = k2; a.tw = time_window; a.act = act; a.n = n;
    return a;
}

int mgb_bundling_decoder_fwd(const float* h, int64_t n, int hidden, const float* u, int ldu, int u_col, const float* w1, const float* b1,
                             int k1, int stride1, const float* w2, const float* b2, int k2, int time_window, int act, const float* dt,
                             float* out, void* stream) {
    return decoder_fwd(dec_args(h, n, u, ldu, u_col, w1, b1, k1, stride1, w2, b2, k2, time_window, act, dt), hidden, out, STREAM(stream));
}

size_t mgb_bundling_decoder_bwd_workspace(int64_t n) { return decoder_bwd_workspace(n); }

int mgb_bundling_decoder_bwd(const float* h, int64_t n, int hidden, const float* u, int ldu, int u_col, const float* w1, const float* b1,
                             int k1, int stride1, const float* w2, const float* b2, int k2, int time_window, int act, const float* dt,
                             const float* dout, float* dh, float* du, int lddu, float* dw1, float* db1, float* dw2, float* db2,
                             void* workspace, size_t workspace_bytes, void* stream) {
    return decoder_bwd(dec_args(h, n, u, ldu, u_col, w1, b1, k1, stride1, w2, b2, k2, time_window, act, dt), hidden, dout, dh, du, lddu,
                       dw1, db1, dw2, db2, workspace, workspace_bytes, STREAM(stream));
}

int mgb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1, double beta2,
                  double eps, double weight_decay, int64_t step, double grad_scale, void* stream) {
    return adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, STREAM(stream));
}

// ---- backward of nn.Linear (+ activation) on the tensor cores: data gradient through the forward's weight images read
// MN-major (linear_tc.cu), weight + bias gradients in one split-K launch (wgrad_tc)
size_t mgb_linear_tc_bwd_workspace(int64_t rows, int in_features, int out_features) {
    if (out_features != 128 || (in_features != 128 && in_features != 256)) return 0;
    return wgrad_tc_workspace(rows) + 1024;
}

int mgb_linear_tc_bwd(const float* dy, const float* y_pre, int act, const float* x, int64_t rows, int in_features, int out_features,
                      const float* packed, float* dx, float* dw, float* db, int accumulate_params, int precision, void* workspace,
                      size_t workspace_bytes, void* stream) {
    MGB_REQUIRE(out_features == 128 && (in_features == 128 || in_features == 256), "linear_tc_bwd: unsupported shape %d -> %d", in_features, out_features);
    MGB_REQUIRE(precision == 1 || precision == 2, "linear_tc_bwd: precision must be 1 (bf16 hi/lo split) or 2 (bf16)");
    MGB_REQUIRE(act >= 0 && act <= 2 && (act == 0 || y_pre != nullptr), "linear_tc_bwd: an activation needs the saved pre-activation");
    MGB_REQUIRE(rows >= 0 && rows < ((int64_t)1 << 31), "linear_tc_bwd: row count out of range");
    if (rows == 0) return MGB_OK;
    const int nk = in_features / 128;
    if (dw) {
        WgradTcArgs w{};
        w.dy = dy; w.lddy = 128; w.ny = 1; w.y_pre = act ? y_pre : nullptr; w.ldyp = 128; w.y_act = act;
        w.nx = nk; w.rows = rows;
        WgradTcOut o[2] = {};
        for (int i = 0; i < nk; ++i) {
            w.x[i] = x + 128 * i; w.ldx[i] = in_features; w.x_act[i] = ACT_NONE;
            o[i].dw = dw + 128 * i; o[i].lddw = in_features; o[i].n_valid = 128; o[i].k_valid = 128; o[i].accumulate = accumulate_params;
        }
        w.db[0] = db; w.db_accumulate = accumulate_params;
        MGB_TRY(launch_wgrad_tc(precision, w, o, workspace, workspace_bytes, STREAM(stream)));
    }
    if (dx) {
        LinTcArgs a{};
        a.src[0] = dy; a.ld[0] = 128; a.nk = 1; a.pre = act ? y_pre : nullptr; a.ldpre = 128; a.pre_act = act;
        a.nm = nk;
        for (int m = 0; m < nk; ++m) a.tile_of[m][0] = m;
        a.act = ACT_NONE; a.y = dx; a.ldy = in_features; a.rows = rows;
        // weights in tensor memory (linear_ts.cu): the W^T blocks sit behind the nk swizzled images of the packed block
        a.wimg = packed + (size_t)nk * 128 * 128;
        MGB_TRY(launch_linear_ts(precision, a, STREAM(stream)));
    }
    return MGB_OK;
}

size_t mgb_linear_bwd_workspace(int64_t rows, int in_features, int out_features) {
    if (small_linear_ok(in_features, out_features)) return small_linear_bwd_workspace(rows, in_features);
    return wgrad_workspace_bytes((int)rows, out_features, in_features) + 1024;
}

int mgb_linear_bwd(const float* dy, const float* y_pre, int act, const float* x, int64_t rows, int in_features,
                   int out_features, const float* w, float* dx, float* dw, float* db, int accumulate_params,
                   void* workspace, size_t workspace_bytes, void* stream) {
    MGB_REQUIRE(rows >= 0 && rows < ((int64_t)1 << 31), "linear_bwd: row count out of range");
    MGB_REQUIRE(act == 0 || y_pre != nullptr || rows == 0, "linear_bwd: an activation needs the saved pre-activation");
    if (small_linear_ok(in_features, out_features))
        return small_linear_bwd(dy, y_pre, act, x, rows, in_features, w, dx, dw, db, accumulate_params, workspace, workspace_bytes,
                                STREAM(stream));
    if (dw) {
        WgradArgs wg{};
        wg.dy = dy; wg.lddy = out_features; wg.y_pre = act ? y_pre : nullptr; wg.y_act = act;
        wg.a.p[0] = x; wg.a.ld[0] = in_features; wg.a.k[0] = in_features; wg.a.nseg = 1;
        wg.rows = (int)rows; wg.N = out_features; wg.K = in_features;
        wg.dw = dw; wg.lddw = in_features; wg.db = db; wg.accumulate = accumulate_params;
        MGB_TRY(launch_wgrad(wg, workspace, workspace_bytes, STREAM(stream)));
    }
    if (dx) {
        GemmArgs g{};
        g.a.p[0] = dy; g.a.ld[0] = out_features; g.a.k[0] = out_features; g.a.nseg = 1;
        g.a.pre = act ? y_pre : nullptr; g.a.pre_act = act;
        g.b = w; g.ldb = in_features; g.c = dx; g.ldc = in_features;
        g.M = (int)rows; g.N = in_features; g.K = out_features;
        MGB_TRY(launch_gemm(g, STREAM(stream)));
    }
    return MGB_OK;
}

int mgb_layernorm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int cols, float* y,
                      float* stats, void* stream) {
    return launch_layernorm_fwd(x, gamma, beta, y, stats, rows, cols, STREAM(stream));
}

int mgb_layernorm_residual_fwd(const float* x, const float* gamma, const float* beta, const float* residual, int64_t rows, int cols,
                               float* y, void* stream) {
    return launch_layernorm_fwd(x, gamma, beta, y, nullptr, rows, cols, STREAM(stream), residual);
}

size_t mgb_layernorm_bwd_workspace(int64_t rows, int cols) { return layernorm_bwd_workspace_bytes(rows, cols); }

int mgb_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* stats, int64_t rows, int cols,
                      float* dx, float* dgamma, float* dbeta, int accumulate_params, void* workspace,
                      size_t workspace_bytes, void* stream) {
    return launch_layernorm_bwd(dy, x, gamma, stats, dx, dgamma, dbeta, accumulate_params, rows, cols, workspace,
                                workspace_bytes, STREAM(stream));
}

size_t mgb_instance_norm_workspace(int n_graphs, int max_nodes_per_graph) {
    return inorm_workspace_bytes(n_graphs, max_nodes_per_graph);
}

int mgb_instance_norm_fwd(const float* x, const int64_t* gptr, int n_graphs, int max_nodes_per_graph, float* y,
                          float* rstd, void* workspace, size_t workspace_bytes, void* stream) {
    return instance_norm_fwd(x, gptr, n_graphs, max_nodes_per_graph, y, rstd, workspace, workspace_bytes, STREAM(stream));
}

int mgb_edge_combine_fwd(const float* p, const float* q, const float* r, const int64_t* edge_index, int64_t n_edges, int act,
                         float* out, void* stream) {
    return edge_combine_fwd(p, q, r, edge_index, n_edges, act, out, STREAM(stream));
}
int mgb_relu_mask(const float* dout, const float* out, int64_t n, float* dz, void* stream) {
    return relu_mask(dout, out, n, dz, STREAM(stream));
}
int mgb_segment_sum_rows(const float* rows, int cols, const int32_t* rowptr, const int32_t* idx, int64_t n_nodes, int mean,
                         float* out, int ld_out, void* stream) {
    return segment_sum_rows(rows, cols, rowptr, idx, n_nodes, mean, out, ld_out, STREAM(stream));
}
int mgb_gather_rows(const float* rows, const int64_t* index, const int32_t* rowptr, int64_t n_edges, float* out, void* stream) {
    return gather_rows(rows, index, rowptr, n_edges, out, STREAM(stream));
}

static InrArgs inr_args(const float* a, const float* xlr, const float* lr_coords, const float* hr_coords, const float* t, int ldt,
                        const float* wsmall, int ldw, const int64_t* idx, int k, int64_t n_query, int nq_per_sample, int L, int T,
                        int d, int mode) {
    InrArgs r;
    r.A = a; r.xlr = xlr; r.lr_coords = lr_coords; r.hr_coords = hr_coords; r.t = t; r.ldt = ldt; r.wsmall = wsmall; r.ldw = ldw;
    r.idx = idx; r.k = k; r.n_query = n_query; r.nq_per_sample = nq_per_sample; r.L = L; r.T = T; r.d = d; r.mode = mode;
    return r;
}
int mgb_inr_decode_fwd(const float* a, const float* xlr, const float* lr_coords, const float* hr_coords, const float* t, int ldt,
                       const float* wsmall, int ldw, const int64_t* idx, int k, int64_t n_query, int nq_per_sample, int L, int T,
                       int d, int mode, float* z, void* stream) {
    return inr_decode_fwd(inr_args(a, xlr, lr_coords, hr_coords, t, ldt, wsmall, ldw, idx, k, n_query, nq_per_sample, L, T, d, mode),
                          z, STREAM(stream));
}
size_t mgb_inr_decode_bwd_workspace(int64_t n_query) { return inr_decode_bwd_workspace(n_query); }
int mgb_inr_decode_bwd(const float* a, const float* xlr, const float* lr_coords, const float* hr_coords, const float* t, int ldt,
                       const float* wsmall, int ldw, const int64_t* idx, int k, int64_t n_query, int nq_per_sample, int L, int T,
                       int d, int mode, const float* dz, float* g, float* sx, float* dwsmall, int accumulate, void* workspace,
                       size_t workspace_bytes, void* stream) {
    return inr_decode_bwd(inr_args(a, xlr, lr_coords, hr_coords, t, ldt, wsmall, ldw, idx, k, n_query, nq_per_sample, L, T, d, mode),
                          dz, g, sx, dwsmall, ldw, accumulate, workspace, workspace_bytes, STREAM(stream));
}

#ifdef MGB_TIMELINE
int mgb_debug_set_timeline(long long* p) { return mgb::set_timeline_buffer(p); }
int mgb_debug_set_ie_timeline(long long* p) { return mgb::set_ie_timeline_buffer(p); }
int mgb_debug_set_ib_timeline(long long* p) { return mgb::set_ib_timeline_buffer(p); }
int mgb_debug_set_lt_timeline(long long* p) { return mgb::set_lt_timeline_buffer(p); }
int mgb_debug_set_nu_timeline(long long* p) { return mgb::set_nu_timeline_buffer(p); }
#endif

int mgb_umma_selftest(const float* a, const float* b, int a_mn_major, int b_mn_major, int lbo_mn, int sbo_mn, float* d,
                      void* stream) {
    return umma_selftest(a, b, a_mn_major, b_mn_major, lbo_mn, sbo_mn, d, STREAM(stream));
}

size_t mgb_sort_workspace(int64_t n) { return sort_workspace_bytes(n); }
int mgb_sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                       int64_t n, int bits, void* workspace, size_t workspace_bytes, void* stream) {
    return radix_sort_pairs(keys_in, vals_in, keys_out, vals_out, n, bits, workspace, workspace_bytes, STREAM(stream));
}
size_t mgb_scan_workspace(int64_t n) { return scan_workspace_bytes(n); }
int mgb_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* workspace, size_t workspace_bytes,
                           void* stream) {
    return exclusive_scan_i32(in, out, n, workspace, workspace_bytes, STREAM(stream));
}

}  // extern "C"
