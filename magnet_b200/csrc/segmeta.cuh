// magnet_b200 — per-tile segment metadata shared by the fused edge kernels (gnn_edge_tc.cu, in_edge_tc.cu):
// the epilogues (thread = channel, columns = destination-sorted edge positions) reduce messages per destination
// in registers with warp-uniform control flow; what they need per tile is prepared ahead of time by one warp.
#pragma once
#include "internal.cuh"

namespace mgb {

constexpr int SEG_H = 128;          // hidden width (row length of the partial / output rows)

// ---- per-tile segment metadata, produced ahead of time by the meta warp ----------------------------------
template <int TE>
struct alignas(16) TileMetaT {
    int dst[TE];           // destination node of every position (-1 past the end of the edge list)
    int src[TE];
    float scale[TE];       // at segment-END positions: factor applied to the segment sum before it is stored
                           // (1/in-degree for a mean over a segment that lies inside the tile, else 1)
    float* out[TE];        // at segment-END positions: row (channel 0) the segment sum goes to — the destination's
                           // output row, or this tile's head / tail partial row when the segment crosses a tile boundary
    uint32_t qoff[TE];     // element offset of Q[src] inside pq: max(src, 0) * 256 + 128
    int segdst[TE];        // the tile's segments in order: destination node, 1 / in-degree
    float seginv[TE];
    uint32_t endmask[TE / 32];   // bit p of the mask: position p is the last of its segment within the tile
    uint32_t flushmask[TE / 32]; // bit p: a running segment sum is stored after position p (segment ends + sub-tile ends)
    int nseg;
};

// one warp; positions p = j*32 + lane
template <int TE>
__device__ __forceinline__ void build_tile_meta(TileMetaT<TE>* M, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ dstv,
                                                const int32_t* __restrict__ srcv, int64_t n_edges, int64_t tile, int lane,
                                                float* out_base, int ld_out, bool mean, float* part_head, float* part_tail,
                                                int flush_te) {
    // flush_te: granularity at which running segment sums are cut and stored; sums cut by a sub-tile
    // boundary go to that sub-tile's head / tail partial row (sub-tile id = e / flush_te) and are merged by the fix-up kernel
    const int64_t e0 = tile * TE;
    const int64_t e1 = (e0 + TE < n_edges) ? e0 + TE : n_edges;
    int base = 0;
#pragma unroll
    for (int j = 0; j < TE / 32; ++j) {
        const int p = j * 32 + lane;
        const int64_t e = e0 + p;
        const bool valid = e < e1;
        const int d = valid ? dstv[e] : -1;
        const int sidx = valid ? srcv[e] : -1;
        const int nxt = (valid && e + 1 < e1) ? dstv[e + 1] : -2;
        const int prv = (valid && p > 0) ? dstv[e - 1] : -2;
        const bool is_end = valid && nxt != d;
        const bool is_start = valid && prv != d;
        const bool is_flush = is_end || (valid && (p % flush_te) == flush_te - 1);
        int s0 = 0, s1 = 1;
        if (is_flush || is_start) {
            s0 = rowptr[d];
            s1 = rowptr[d + 1];
        }
        const float inv = 1.0f / (float)(s1 - s0);
        M->dst[p] = d;
        M->src[p] = sidx;
        M->qoff[p] = (uint32_t)(sidx < 0 ? 0 : sidx) * (2u * SEG_H) + SEG_H;
        const int64_t f0 = e0 + (p / flush_te) * flush_te;                 // bounds of this position's sub-tile
        const int64_t f1 = (f0 + flush_te < e1) ? f0 + flush_te : e1;
        const bool inside = (int64_t)s0 >= f0 && (int64_t)s1 <= f1;
        M->scale[p] = (inside && mean) ? inv : 1.0f;
        if (is_flush)
            M->out[p] = inside ? out_base + (int64_t)d * ld_out : ((int64_t)s0 < f0 ? part_head : part_tail) + (f0 / flush_te) * SEG_H;
        const uint32_t em = __ballot_sync(0xffffffffu, is_end);
        const uint32_t fm = __ballot_sync(0xffffffffu, is_flush);
        const uint32_t sm = __ballot_sync(0xffffffffu, is_start);
        if (lane == 0) { M->endmask[j] = em; M->flushmask[j] = fm; }
        if (is_start) {
            const int idx = base + __popc(sm & ((1u << lane) - 1u));
            M->segdst[idx] = d;
            M->seginv[idx] = inv;
        }
        base += __popc(sm);
    }
    if (lane == 0) M->nseg = base;
}


// merges the head / tail partial rows of segments cut by a sub-tile boundary (gnn_edge_tc.cu)
int launch_segment_fixup(const int32_t* rowptr, const int32_t* dstv, int64_t n_edges, int te, const float* part_head,
                         const float* part_tail, float* out, int ld_out, int mean, cudaStream_t s);

}  // namespace mgb
