// magnet_b200 — grid-hashed radius-graph / kNN builders and CSR aggregation plans (sm_100a).
//
// Replaces, bit-exactly, the brute-force torch_cluster searches the reference calls
// (radius_graph: models/mpnn_2d.py:245, models/mpnn.py:245, models/magnet_gnn.py:293;
//  knn: models/magnet_gnn.py:247).  Semantics that must be reproduced:
//   * radius: per centre, the FIRST `max` in-radius nodes IN NODE-INDEX ORDER (not the nearest),
//     strict `dist < r*r`, fp32 `dist = fma(d_k, d_k, dist)` in dimension order, r*r evaluated
//     in double then rounded to fp32; loop=False searches max+1 hits and then drops the centre.
//   * knn: k smallest (dist, index) pairs, ascending.
//
// Design: points are binned into a uniform grid (cell edge >= 1.01 r), sorted by cell with the
// stable radix sort (so each cell's list is ascending in node index), and every centre walks a
// 3^d-way merge of its neighbour cells' lists in node-index order with early exit after `max`
// hits — the reference's scan restricted to a superset of the in-radius nodes, so the cost is
// O(max / hit-rate) per centre whatever the local density.  Threads are assigned to centres in
// cell order, so a warp shares its candidate cells (L1-resident broadcast loads).  All integer /
// compare work, HBM- and latency-bound: no tensor cores here.
#include "common.cuh"
#include "grid.cuh"
#include <math.h>

namespace mgb {

__device__ __forceinline__ int float_order_key(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float float_from_order_key(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void bbox_init_kernel(int* bbox) {
    if (threadIdx.x < 2) bbox[threadIdx.x] = 0x7fffffff;            // mins
    else if (threadIdx.x < 4) bbox[threadIdx.x] = (int)0x80000000;  // maxs
}

__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ pos, int64_t n, int d, int* __restrict__ bbox) {
    int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = (int)0x80000000, mxy = (int)0x80000000;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int kx = float_order_key(pos[i * d]);
        int ky = d > 1 ? float_order_key(pos[i * d + 1]) : float_order_key(0.0f);
        mnx = min(mnx, kx); mxx = max(mxx, kx);
        mny = min(mny, ky); mxy = max(mxy, ky);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bbox[0], mnx); atomicMin(&bbox[1], mny);
        atomicMax(&bbox[2], mxx); atomicMax(&bbox[3], mxy);
    }
}

// h_min: smallest admissible cell edge (1.01 r for radius search; 0 => pick ~target_ppc points per cell)
__global__ void grid_params_kernel(const int* __restrict__ bbox, int d, float h_min, float target_ppc, int64_t n,
                                   int n_samples, int64_t max_cells, GridParams* __restrict__ gp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float mnx = float_from_order_key(bbox[0]), mny = float_from_order_key(bbox[1]);
    float mxx = float_from_order_key(bbox[2]), mxy = float_from_order_key(bbox[3]);
    if (n == 0) { mnx = mny = 0.f; mxx = mxy = 0.f; }
    float ex = fmaxf(mxx - mnx, 0.f), ey = d > 1 ? fmaxf(mxy - mny, 0.f) : 0.f;
    const int CAP = 2048;                       // cells per dimension: keeps fp error of the cell index << margin
    float h = h_min;
    if (h <= 0.f) {                             // density-driven cell size (kNN)
        float per_sample = fmaxf((float)n / (float)max(n_samples, 1), 1.f);
        if (d > 1) h = sqrtf(fmaxf(ex * ey, 1e-30f) * target_ppc / per_sample);
        else h = fmaxf(ex, 1e-30f) * target_ppc / per_sample;
    }
    h = fmaxf(h, 1e-30f);
    float hx = fmaxf(h, ex / (float)CAP * 1.001f), hy = fmaxf(h, ey / (float)CAP * 1.001f);
    int ncx = 1, ncy = 1;
    for (int it = 0; it < 64; ++it) {
        ncx = min(CAP, (int)floorf(ex / hx) + 1);
        ncy = d > 1 ? min(CAP, (int)floorf(ey / hy) + 1) : 1;
        if ((int64_t)ncx * ncy * n_samples <= max_cells) break;
        hx *= 1.25f; hy *= 1.25f;
    }
    if ((int64_t)ncx * ncy * n_samples > max_cells) { ncx = 1; ncy = 1; }
    gp->min_x = mnx; gp->min_y = mny;
    gp->hx = hx; gp->hy = hy;
    gp->inv_hx = 1.0f / hx; gp->inv_hy = 1.0f / hy;
    gp->ncx = ncx; gp->ncy = ncy;
    gp->cells_per_sample = ncx * ncy;
}

__global__ void __launch_bounds__(256)
cell_assign_kernel(const float* __restrict__ pos, int64_t n, int d, const int64_t* __restrict__ ptr, int n_samples,
                   const GridParams* __restrict__ gpp, uint32_t* __restrict__ cell, uint32_t* __restrict__ ident) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GridParams gp = *gpp;
    float x = pos[i * d], y = d > 1 ? pos[i * d + 1] : 0.f;
    int cx, cy;
    cell_coords(gp, x, y, cx, cy);
    int b = sample_of(ptr, n_samples, i);
    cell[i] = (uint32_t)(b * gp.cells_per_sample + cy * gp.ncx + cx);
    ident[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
cell_pack_kernel(const float* __restrict__ pos, int64_t n, int d, const uint32_t* __restrict__ sorted_cell,
                 const uint32_t* __restrict__ sorted_idx, CellPoint* __restrict__ pts) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int i = (int)sorted_idx[p];
    CellPoint c;
    c.x = pos[(int64_t)i * d];
    c.y = d > 1 ? pos[(int64_t)i * d + 1] : 0.f;
    c.idx = i;
    c.cell = (int)sorted_cell[p];
    pts[p] = c;
}

// ------------------------------------------------------------------------------------------
// radius search: NL-way merge (NL = 3 or 9) of cell lists in node-index order with early exit
// ------------------------------------------------------------------------------------------
template <int NL>
__global__ void __launch_bounds__(128)
radius_search_kernel(const CellPoint* __restrict__ pts, const int32_t* __restrict__ cell_start, int64_t n,
                     const GridParams* __restrict__ gpp, float r2, int max_hits, int drop_self, int cap,
                     int32_t* __restrict__ nbr /*[n][cap]*/, int32_t* __restrict__ deg /*[n]*/) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const GridParams gp = *gpp;
    const CellPoint me = pts[p];
    const int local = me.cell % gp.cells_per_sample;
    const int sample_base = me.cell - local;
    const int cy = local / gp.ncx, cx = local - cy * gp.ncx;

    int cur[NL], end[NL], head[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        int dx = (l % 3) - 1, dy = (NL == 9) ? (l / 3) - 1 : 0;
        int nx = cx + dx, ny = cy + dy;
        bool ok = nx >= 0 && nx < gp.ncx && ny >= 0 && ny < gp.ncy;
        int c = sample_base + ny * gp.ncx + nx;
        cur[l] = ok ? cell_start[c] : 0;
        end[l] = ok ? cell_start[c + 1] : 0;
        head[l] = (cur[l] < end[l]) ? pts[cur[l]].idx : 0x7fffffff;
    }
    int32_t* out = nbr + (int64_t)me.idx * cap;
    int hits = 0, written = 0;
    while (hits < max_hits) {
        int best = 0x7fffffff, bl = -1;
#pragma unroll
        for (int l = 0; l < NL; ++l)
            if (head[l] < best) { best = head[l]; bl = l; }
        if (bl < 0) break;
        int at = 0;
#pragma unroll
        for (int l = 0; l < NL; ++l)
            if (l == bl) {
                at = cur[l];
                cur[l] = at + 1;
                head[l] = (at + 1 < end[l]) ? pts[at + 1].idx : 0x7fffffff;
            }
        const CellPoint q = pts[at];
        // reference arithmetic: dist = 0; dist = fma(dx,dx,dist); dist = fma(dy,dy,dist)  (x[n_x] - y[n_y])
        float ddx = q.x - me.x;
        float dist = __fmaf_rn(ddx, ddx, 0.0f);
        if (NL == 9) { float ddy = q.y - me.y; dist = __fmaf_rn(ddy, ddy, dist); }
        if (dist < r2) {
            ++hits;
            if (!(drop_self && q.idx == me.idx)) { if (written < cap) out[written] = q.idx; ++written; }
        }
    }
    deg[me.idx] = min(written, cap);
}

__global__ void __launch_bounds__(256)
radius_emit_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ rowptr, int64_t n, int cap,
                   int centre_row, int64_t n_edges, int64_t* __restrict__ edge_index /*[2][E]*/,
                   int32_t* __restrict__ col /*[E] or null*/) {
    // one warp per centre: up to 33 contiguous outputs
    int64_t centre = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (centre >= n) return;
    int s = rowptr[centre], e = rowptr[centre + 1];
    for (int k = lane; k < e - s; k += 32) {
        int v = nbr[centre * cap + k];
        edge_index[(int64_t)centre_row * n_edges + s + k] = centre;
        edge_index[(int64_t)(1 - centre_row) * n_edges + s + k] = v;
        if (col) col[s + k] = v;
    }
}

static int build_grid(const float* pos, int64_t n, int d, const int64_t* ptr, int n_samples, float h_min,
                      float target_ppc, int64_t max_cells, Workspace& ws, GridParams** gp_out, CellPoint** pts_out,
                      int32_t** cell_start_out, cudaStream_t s, bool reuse = false) {
    int* bbox = ws.take<int>(4);
    GridParams* gp = ws.take<GridParams>(1);
    uint32_t* cell = ws.take<uint32_t>(n);
    uint32_t* ident = ws.take<uint32_t>(n);
    uint32_t* scell = ws.take<uint32_t>(n);
    uint32_t* sidx = ws.take<uint32_t>(n);
    CellPoint* pts = ws.take<CellPoint>(n);
    int32_t* cell_start = ws.take<int32_t>(max_cells + 2);
    size_t sort_bytes = sort_workspace_bytes(n);
    char* sort_ws = ws.take<char>(sort_bytes);
    MGB_WS_CHECK(ws);
    *gp_out = gp; *pts_out = pts; *cell_start_out = cell_start;
    if (reuse) return MGB_OK;            // the workspace still holds the grid of an earlier call over the same points
    bbox_init_kernel<<<1, 32, 0, s>>>(bbox);
    MGB_LAUNCH_CHECK();
    if (n > 0) {
        int64_t nb64 = ceil_div<int64_t>(n, 256); int blocks = (int)(nb64 < 4 * sm_count() ? nb64 : 4 * sm_count());
        bbox_kernel<<<blocks, 256, 0, s>>>(pos, n, d, bbox);
        MGB_LAUNCH_CHECK();
    }
    grid_params_kernel<<<1, 32, 0, s>>>(bbox, d, h_min, target_ppc, n, n_samples, max_cells, gp);
    MGB_LAUNCH_CHECK();
    if (n > 0) {
        cell_assign_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, s>>>(pos, n, d, ptr, n_samples, gp, cell, ident);
        MGB_LAUNCH_CHECK();
        int bits = 1;
        while (((int64_t)1 << bits) < max_cells + 1) ++bits;
        MGB_TRY(radix_sort_pairs(cell, ident, scell, sidx, n, bits, sort_ws, sort_bytes, s));
        cell_pack_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, s>>>(pos, n, d, scell, sidx, pts);
        MGB_LAUNCH_CHECK();
    }
    MGB_TRY(segment_starts(scell, n, cell_start, max_cells + 1, s));
    *gp_out = gp; *pts_out = pts; *cell_start_out = cell_start;
    return MGB_OK;
}

static int64_t default_max_cells(int64_t n, int n_samples) {
    int64_t m = 2 * n + 64 * (int64_t)n_samples + 1024;
    if (m > ((int64_t)1 << 26)) m = (int64_t)1 << 26;
    if (m < n_samples) m = n_samples;
    return m;
}

size_t radius_workspace_bytes(int64_t n, int n_samples) {
    int64_t mc = default_max_cells(n, n_samples);
    size_t b = 0;
    b += align_up(16) + align_up(sizeof(GridParams)) + 4 * align_up((size_t)n * 4) + align_up((size_t)n * sizeof(CellPoint));
    b += align_up((size_t)(mc + 2) * 4) + align_up(sort_workspace_bytes(n)) + scan_workspace_bytes(n) + 4096;
    return b;
}

// Phase 1: neighbour lists (padded) + rowptr + edge count (device).  No host sync.
int radius_search(const float* pos, int64_t n, int d, const int64_t* ptr, int n_samples, double r, int max_num_neighbors,
                  int loop, int32_t* nbr, int32_t* deg, int32_t* rowptr, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(d == 1 || d == 2, "radius_graph: only 1-D and 2-D coordinates are supported (got d=%d)", d);
    MGB_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) / 40, "radius_graph: node count out of range");
    MGB_REQUIRE(n_samples >= 1 && max_num_neighbors >= 1, "radius_graph: bad n_samples / max_num_neighbors");
    MGB_REQUIRE(r > 0, "radius_graph: radius must be positive");
    Workspace ws(ws_ptr, ws_bytes);
    const float r2 = (float)(r * r);                       // double product, then rounded (torch_cluster launch)
    const int max_hits = loop ? max_num_neighbors : max_num_neighbors + 1;
    const int cap = max_hits;                              // loop=False can keep max+1 when the centre is beyond slot max+1
    const int64_t mc = default_max_cells(n, n_samples);
    GridParams* gp; CellPoint* pts; int32_t* cell_start;
    MGB_TRY(build_grid(pos, n, d, ptr, n_samples, (float)(r * 1.01), 0.f, mc, ws, &gp, &pts, &cell_start, s));
    if (n > 0) {
        unsigned blocks = (unsigned)ceil_div<int64_t>(n, 128);
        if (d == 2) radius_search_kernel<9><<<blocks, 128, 0, s>>>(pts, cell_start, n, gp, r2, max_hits, loop ? 0 : 1, cap, nbr, deg);
        else        radius_search_kernel<3><<<blocks, 128, 0, s>>>(pts, cell_start, n, gp, r2, max_hits, loop ? 0 : 1, cap, nbr, deg);
        MGB_LAUNCH_CHECK();
    }
    MGB_TRY(exclusive_scan_i32(deg, rowptr, n, ws.base + ws.off, ws.cap > ws.off ? ws.cap - ws.off : 0, s));
    return MGB_OK;
}

int radius_emit(const int32_t* nbr, const int32_t* rowptr, int64_t n, int cap, int centre_row, int64_t n_edges,
                int64_t* edge_index, int32_t* col, cudaStream_t s) {
    MGB_REQUIRE(centre_row == 0 || centre_row == 1, "radius_emit: centre_row must be 0 or 1");
    if (n == 0 || n_edges == 0) return MGB_OK;
    radius_emit_kernel<<<(unsigned)ceil_div<int64_t>(n * 32, 256), 256, 0, s>>>(nbr, rowptr, n, cap, centre_row, n_edges, edge_index, col);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------
// kNN: expanding-ring search over the grid of the x points, per-thread sorted (dist, idx) list
// ------------------------------------------------------------------------------------------
template <int K, int D>
__global__ void __launch_bounds__(128)
knn_kernel(const CellPoint* __restrict__ pts, const int32_t* __restrict__ cell_start, const GridParams* __restrict__ gpp,
           const float* __restrict__ y, int64_t ny, const int64_t* __restrict__ ptr_y, const int64_t* __restrict__ ptr_x,
           int n_samples, int k, int64_t* __restrict__ out_idx /*[ny][k]*/, float* __restrict__ out_dist /*opt*/) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ny) return;
    const GridParams gp = *gpp;
    const float qx = y[q * D], qy = D > 1 ? y[q * D + 1] : 0.f;
    const int b = sample_of(ptr_y, n_samples, q);
    BestList<K> best;
    knn_query<K, D>(pts, cell_start, gp, qx, qy, b, ptr_x[b + 1] - ptr_x[b], k, best);
#pragma unroll
    for (int e = 0; e < K; ++e)
        if (e < k) {
            out_idx[q * k + e] = best.i[e];
            if (out_dist) out_dist[q * k + e] = best.d[e];
        }
}

size_t knn_workspace_bytes(int64_t nx, int n_samples) { return radius_workspace_bytes(nx, n_samples); }

size_t grid_workspace_bytes(int64_t n, int n_samples) { return radius_workspace_bytes(n, n_samples); }

int build_grid_ws(const float* pos, int64_t n, int d, const int64_t* ptr, int n_samples, float h_min, float target_ppc, void* ws_ptr,
                  size_t ws_bytes, GridParams** gp, CellPoint** pts, int32_t** cell_start, cudaStream_t s, int reuse) {
    Workspace ws(ws_ptr, ws_bytes);
    return build_grid(pos, n, d, ptr, n_samples, h_min, target_ppc, default_max_cells(n, n_samples), ws, gp, pts, cell_start, s, reuse != 0);
}

int knn_search(const float* x, int64_t nx, const float* y, int64_t ny, int d, const int64_t* ptr_x, const int64_t* ptr_y,
               int n_samples, int k, int64_t* out_idx, float* out_dist, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(d == 1 || d == 2, "knn: only 1-D and 2-D coordinates are supported (got d=%d)", d);
    MGB_REQUIRE(k >= 1 && k <= 64, "knn: k must be in [1, 64] (got %d)", k);
    MGB_REQUIRE(nx >= 0 && nx < ((int64_t)1 << 31) && ny >= 0, "knn: sizes out of range");
    Workspace ws(ws_ptr, ws_bytes);
    const int64_t mc = default_max_cells(nx, n_samples);
    GridParams* gp; CellPoint* pts; int32_t* cell_start;
    MGB_TRY(build_grid(x, nx, d, ptr_x, n_samples, 0.f, 2.0f, mc, ws, &gp, &pts, &cell_start, s));
    if (ny == 0) return MGB_OK;
    unsigned blocks = (unsigned)ceil_div<int64_t>(ny, 128);
#define MGB_KNN_LAUNCH(KK)                                                                                     \
    do {                                                                                                       \
        if (d == 2) knn_kernel<KK, 2><<<blocks, 128, 0, s>>>(pts, cell_start, gp, y, ny, ptr_y, ptr_x, n_samples, k, out_idx, out_dist); \
        else        knn_kernel<KK, 1><<<blocks, 128, 0, s>>>(pts, cell_start, gp, y, ny, ptr_y, ptr_x, n_samples, k, out_idx, out_dist); \
    } while (0)
    if (k <= 4) MGB_KNN_LAUNCH(4);
    else if (k <= 8) MGB_KNN_LAUNCH(8);
    else if (k <= 16) MGB_KNN_LAUNCH(16);
    else if (k <= 32) MGB_KNN_LAUNCH(32);
    else MGB_KNN_LAUNCH(64);
#undef MGB_KNN_LAUNCH
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------
// CSR aggregation plan from a COO endpoint column (stable: edges of one node keep COO order)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
plan_keys_kernel(const int64_t* __restrict__ key64, int64_t n_edges, int64_t n_nodes, uint32_t* __restrict__ key,
                 uint32_t* __restrict__ ident, int* __restrict__ bad) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    int64_t k = key64[e];
    if (k < 0 || k >= n_nodes) { *bad = 1; k = 0; }
    key[e] = (uint32_t)k;
    ident[e] = (uint32_t)e;
}

__global__ void __launch_bounds__(256)
plan_gather_kernel(const uint32_t* __restrict__ perm, const int64_t* __restrict__ a64, const int64_t* __restrict__ b64,
                   int64_t n_edges, int32_t* __restrict__ a_sorted, int32_t* __restrict__ b_sorted, int32_t* __restrict__ inv_perm) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_edges) return;
    uint32_t e = perm[p];
    if (a_sorted) a_sorted[p] = (int32_t)a64[e];
    if (b_sorted) b_sorted[p] = (int32_t)b64[e];
    if (inv_perm) inv_perm[e] = (int32_t)p;
}

__global__ void __launch_bounds__(256)
plan_compose_kernel(const uint32_t* __restrict__ perm_t, const int32_t* __restrict__ inv_perm, int64_t n_edges,
                    int32_t* __restrict__ pos_t) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_edges) return;
    pos_t[q] = inv_perm[perm_t[q]];
}

size_t csr_plan_workspace_bytes(int64_t n_edges) {
    return 5 * align_up((size_t)(n_edges > 0 ? n_edges : 1) * 4) + align_up(sort_workspace_bytes(n_edges)) + 4096;
}

// agg = edge_index row the messages are reduced at, other = the opposite row.  Outputs (all int32):
//   rowptr[n_nodes+1], perm[E] (COO edge id at each aggregation-order position), dst[E], src[E] (endpoints
//   in aggregation order), and the transposed plan rowptr_t[n_nodes+1], pos_t[E] (aggregation-order positions
//   grouped by `other`, used by the backward pass to reduce dx_j without atomics).
int csr_plan(const int64_t* agg, const int64_t* other, int64_t n_edges, int64_t n_nodes, int32_t* rowptr, int32_t* perm,
             int32_t* dst, int32_t* src, int32_t* rowptr_t, int32_t* pos_t, int* bad_flag, void* ws_ptr, size_t ws_bytes,
             cudaStream_t s) {
    MGB_REQUIRE(n_edges >= 0 && n_edges < ((int64_t)1 << 31) && n_nodes >= 0 && n_nodes < ((int64_t)1 << 31),
                "csr_plan: sizes out of range");
    Workspace ws(ws_ptr, ws_bytes);
    int64_t ne = n_edges > 0 ? n_edges : 1;
    uint32_t* key = ws.take<uint32_t>(ne);
    uint32_t* ident = ws.take<uint32_t>(ne);
    uint32_t* skey = ws.take<uint32_t>(ne);
    uint32_t* sperm = ws.take<uint32_t>(ne);
    int32_t* inv = ws.take<int32_t>(ne);
    size_t sort_bytes = sort_workspace_bytes(n_edges);
    char* sort_ws = ws.take<char>(sort_bytes);
    MGB_WS_CHECK(ws);
    int bits = 1;
    while (((int64_t)1 << bits) < n_nodes) ++bits;
    unsigned blocks = (unsigned)ceil_div<int64_t>(ne, 256);
    MGB_CUDA(cudaMemsetAsync(bad_flag, 0, sizeof(int), s));
    if (n_edges > 0) {
        plan_keys_kernel<<<blocks, 256, 0, s>>>(agg, n_edges, n_nodes, key, ident, bad_flag);
        MGB_LAUNCH_CHECK();
        MGB_TRY(radix_sort_pairs(key, ident, skey, (uint32_t*)perm, n_edges, bits, sort_ws, sort_bytes, s));
        plan_gather_kernel<<<blocks, 256, 0, s>>>((const uint32_t*)perm, agg, other, n_edges, dst, src, inv);
        MGB_LAUNCH_CHECK();
    }
    MGB_TRY(segment_starts(skey, n_edges, rowptr, n_nodes, s));
    if (rowptr_t && pos_t) {
        if (n_edges > 0) {
            plan_keys_kernel<<<blocks, 256, 0, s>>>(other, n_edges, n_nodes, key, ident, bad_flag);
            MGB_LAUNCH_CHECK();
            MGB_TRY(radix_sort_pairs(key, ident, skey, sperm, n_edges, bits, sort_ws, sort_bytes, s));
            plan_compose_kernel<<<blocks, 256, 0, s>>>(sperm, inv, n_edges, pos_t);
            MGB_LAUNCH_CHECK();
        }
        MGB_TRY(segment_starts(skey, n_edges, rowptr_t, n_nodes, s));
    }
    return MGB_OK;
}

}  // namespace mgb
