// magnet_b200 — uniform-grid spatial hash shared by the graph builders (graph.cu) and the fused INR decoder.
#pragma once
#include "common.cuh"

namespace mgb {

struct GridParams {      // written by one device thread, read by every later kernel (no host sync)
    float min_x, min_y;
    float inv_hx, inv_hy;
    float hx, hy;
    int ncx, ncy;
    int cells_per_sample;
};

struct __align__(16) CellPoint { float x, y; int idx; int cell; };

__device__ __forceinline__ int sample_of(const int64_t* __restrict__ ptr, int n_samples, int64_t i) {
    int lo = 0, hi = n_samples;                 // largest b with ptr[b] <= i
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (ptr[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void cell_coords(const GridParams& gp, float x, float y, int& cx, int& cy) {
    cx = (int)floorf((x - gp.min_x) * gp.inv_hx);
    cy = (int)floorf((y - gp.min_y) * gp.inv_hy);
    cx = max(0, min(cx, gp.ncx - 1));
    cy = max(0, min(cy, gp.ncy - 1));
}

template <int K>
struct BestList {
    float d[K];
    int i[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int e = 0; e < K; ++e) { d[e] = 1e10f; i[e] = -1; }
    }
    // keeps the K smallest (dist, idx) pairs in ascending order; equal to the reference's stable
    // insertion over an ascending-index scan (ties: lower index first).
    __device__ __forceinline__ void push(float dist, int idx, int k) {
        float cd = dist; int ci = idx;
#pragma unroll
        for (int e = 0; e < K; ++e) {
            if (e < k) {
                // lexicographic (dist, idx); empty slots are (1e10, -1) and lose to any dist < 1e10
                bool better = (cd < d[e]) || (cd == d[e] && ci < i[e]);
                if (better) { float td = d[e]; int ti = i[e]; d[e] = cd; i[e] = ci; cd = td; ci = ti; }
            }
        }
    }
};

// One query of the expanding-ring kNN over the grid of the x points: the k (<= K) smallest (dist, idx) pairs of sample b in
// ascending order, ties -> lower index (torch_cluster.knn's order; models/magnet_gnn.py:247).  Shared by knn_kernel
// (graph.cu) and the fused INR decoder (mlp_chain_tc.cu), whose search warps run it in place.
template <int K, int D>
__device__ __forceinline__ void knn_query(const CellPoint* __restrict__ pts, const int32_t* __restrict__ cell_start, const GridParams& gp,
                                          float qx, float qy, int b, int64_t n_in_sample, int k, BestList<K>& best) {
    const int sample_base = b * gp.cells_per_sample;
    int cx, cy;
    cell_coords(gp, qx, qy, cx, cy);
    best.init();
    const int want = (int)(n_in_sample < (int64_t)k ? n_in_sample : (int64_t)k);
    int found = 0;
    const int max_ring = max(gp.ncx, gp.ncy);
    for (int ring = 0; ring <= max_ring && want > 0; ++ring) {
        const int x0 = cx - ring, x1 = cx + ring, y0 = D > 1 ? cy - ring : 0, y1 = D > 1 ? cy + ring : 0;
        for (int yy = max(y0, 0); yy <= min(y1, gp.ncy - 1); ++yy) {
            const bool edge_row = (D > 1) && (yy == y0 || yy == y1);
            const int step = edge_row ? 1 : max(x1 - x0, 1);      // interior rows: only the two end columns
            for (int xx = x0; xx <= x1; xx += step) {
                if (xx < 0 || xx >= gp.ncx) continue;
                const int c = sample_base + yy * gp.ncx + xx;
                const int s = cell_start[c], e = cell_start[c + 1];
                for (int p = s; p < e; ++p) {
                    const CellPoint c4 = pts[p];
                    float ddx = c4.x - qx;
                    float dist = __fmaf_rn(ddx, ddx, 0.0f);
                    if (D > 1) { float ddy = c4.y - qy; dist = __fmaf_rn(ddy, ddy, dist); }
                    ++found;
                    best.push(dist, c4.idx, k);
                }
                if (ring == 0) break;
            }
        }
        if (found >= want) {
            // distance from the query to the nearest face of the visited block that still has grid beyond it
            float bd = 3.0e38f;
            const float m = 2e-3f;    // cell-index rounding margin, in cell units
            if (x0 > 0)          bd = fminf(bd, qx - (gp.min_x + ((float)x0 + m) * gp.hx));
            if (x1 < gp.ncx - 1) bd = fminf(bd, (gp.min_x + ((float)(x1 + 1) - m) * gp.hx) - qx);
            if (D > 1) {
                if (y0 > 0)          bd = fminf(bd, qy - (gp.min_y + ((float)y0 + m) * gp.hy));
                if (y1 < gp.ncy - 1) bd = fminf(bd, (gp.min_y + ((float)(y1 + 1) - m) * gp.hy) - qy);
            }
            if (bd >= 3.0e38f) break;                              // the block covers the sample's whole grid
            float kth = 0.f;
#pragma unroll
            for (int e = 0; e < K; ++e) if (e == want - 1) kth = best.d[e];
            if (bd > 0.f && kth < bd * bd * 0.9999f) break;      // strict: an outside point can neither beat nor tie
        }
    }
}

// graph.cu: bins the points of every sample into the grid (cell list sorted by cell, ascending node index inside a cell).
// h_min > 0: cell edge >= h_min (radius search); h_min == 0: about target_ppc points per cell (kNN).
size_t grid_workspace_bytes(int64_t n, int n_samples);
int build_grid_ws(const float* pos, int64_t n, int d, const int64_t* ptr, int n_samples, float h_min, float target_ppc, void* ws,
                  size_t ws_bytes, GridParams** gp, CellPoint** pts, int32_t** cell_start, cudaStream_t s, int reuse = 0);

}  // namespace mgb
