// magnet_b200 — fused edge kernels of GNN_Layer on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Reference: GNN_Layer.message (models/mpnn_2d.py:73-79) + mean aggregation (:46) and their backward.
// With the first Linear factorised per node (gnn_layer.cu), the per-edge work is
//     h1_e = Swish(P[dst_e] + Q[src_e]),   m_e = Swish(W2 h1_e + b2),   agg[i] = mean_{e->i} m_e.
//
// The 128x128 contraction runs TRANSPOSED on the tensor core:  D^T[n][e] = sum_k W2[n][k] h1[e][k]
//   A = W2 (M = output channel n), resident in shared memory for the whole persistent CTA,
//   B = the gathered/activated edge tile (N = 128 edge positions of the dst-sorted order),
//   D^T in TMEM: lane = channel, column = edge position.
// So an epilogue thread owns ONE channel and walks the 128 edge columns in order: bias + Swish +
// the segmented mean over the destination-sorted positions happen in registers with warp-uniform
// control flow — no atomics, no shuffles, no shared-memory staging — and every flush is a coalesced
// 128-byte store per warp.
//
// Warp roles (416 threads, one persistent CTA per SM):
//   warps 0-3   epilogue   TMEM -> registers -> bias/Swish -> segmented mean -> agg / boundary partials
//   warp  4     MMA issue  (one elected lane), TMEM allocation
//   warps 5-12  producers  coalesced float4 gathers of P[dst], Q[src] (P reused along a segment),
//                          Swish, bf16 (hi[/lo]) conversion, swizzled K-major tile stores
// Pipelines: 2 smem stages (full/empty mbarriers) x 2 TMEM accumulator stages (tfull/tempty).
// Precision: NSPLIT = 1 -> plain bf16 operands (1e-2 contract); NSPLIT = 2 -> hi/lo bf16 split of both
// operands, three MMAs (hi*hi + hi*lo + lo*hi), fp32 accumulation: error ~2^-17, inside the 1e-5 contract.
#include "internal.cuh"
#include "umma.cuh"
#include "tc_common.cuh"
#include "segmeta.cuh"

namespace mgb {

constexpr int TCH = 128;            // hidden width
constexpr int TCE = 128;            // edge positions per tile (MMA N)
constexpr int TC_STAGES = 2;
constexpr int TC_EPI_WARPS = 4, TC_PROD_WARPS = 8;
constexpr int TC_MMA_WARP = TC_EPI_WARPS, TC_META_WARP = TC_EPI_WARPS + 1, TC_PROD_WARP0 = TC_EPI_WARPS + 2;
constexpr int TC_THREADS = (TC_EPI_WARPS + 2 + TC_PROD_WARPS) * 32;   // 448
constexpr int META_STAGES = 4;

// ---- W2 -> swizzled bf16 images (hi, lo) ----------------------------------------------------------
__global__ void pack_w2_image_kernel(const float* __restrict__ W2, unsigned char* __restrict__ img) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= TCH * TCH) return;
    const int n = idx >> 7, k = idx & 127;
    __nv_bfloat16 hi, lo;
    umma::split_bf16(W2[idx], hi, lo);
    const uint32_t off = umma::tile_off(128, n, k);
    *reinterpret_cast<__nv_bfloat16*>(img + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(img + TILE_BYTES + off) = lo;
}

int pack_w2_image(const float* W2, void* img, cudaStream_t s) {
    pack_w2_image_kernel<<<TCH * TCH / 256, 256, 0, s>>>(W2, (unsigned char*)img);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// Segments cut by a sub-tile boundary: merge the tail partial of the sub-tile the segment starts in, the head partials
// of the sub-tiles it covers and the head partial of the sub-tile it ends in.  One warp per sub-tile (4 channels per lane).
__global__ void __launch_bounds__(256)
segment_fixup_tc_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ dstv, int64_t n_edges, int te,
                        const float* __restrict__ part_head, const float* __restrict__ part_tail, float* __restrict__ out,
                        int ld_out, int mean) {
    const int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5) + 1;
    const int64_t e0 = t * te;
    if (e0 >= n_edges) return;
    const int node = dstv[e0];
    const int64_t s0 = rowptr[node], s1 = rowptr[node + 1];
    if (!(s0 < e0 && s1 <= e0 + te)) return;
    const int c = (threadIdx.x & 31) * 4;
    const int64_t t0 = s0 / te;
    float4 acc = *reinterpret_cast<const float4*>(part_tail + t0 * TCH + c);
    for (int64_t tt = t0 + 1; tt <= t; ++tt) {
        const float4 h = *reinterpret_cast<const float4*>(part_head + tt * TCH + c);
        acc.x += h.x; acc.y += h.y; acc.z += h.z; acc.w += h.w;
    }
    if (mean) {
        const float inv = 1.0f / (float)(s1 - s0);       // same rounding as the in-tile mean (multiply by 1/deg)
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    }
    *reinterpret_cast<float4*>(out + (int64_t)node * ld_out + c) = acc;
}

using TileMeta = TileMetaT<TCE>;

int launch_segment_fixup(const int32_t* rowptr, const int32_t* dstv, int64_t n_edges, int te, const float* part_head,
                         const float* part_tail, float* out, int ld_out, int mean, cudaStream_t s) {
    const int64_t subtiles = ceil_div<int64_t>(n_edges, te);
    if (subtiles <= 1) return MGB_OK;
    segment_fixup_tc_kernel<<<(unsigned)ceil_div<int64_t>(subtiles - 1, 8), 256, 0, s>>>(rowptr, dstv, n_edges, te, part_head, part_tail, out, ld_out, mean);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ==================================================================================================
// Forward
// ==================================================================================================
struct EdgeFwdTcArgs {
    const float* pq;         // [N][256]  P | Q  (fp32)
    const int32_t* rowptr;   // [N+1]
    const int32_t* dstv;     // [E]
    const int32_t* srcv;     // [E]
    int64_t n_edges;
    const unsigned char* w2img;   // swizzled bf16 images of W2: hi | lo
    const float* b2;
    float* agg;              // [N][128] pre-zeroed
    float* part_head;        // [tiles][128]
    float* part_tail;
};

template <int NSPLIT>
constexpr size_t edge_fwd_tc_smem() {
    return 1024 + (size_t)NSPLIT * TILE_BYTES * (1 + TC_STAGES) + META_STAGES * sizeof(TileMeta) + 256;
}

// forward warp roles (whole warpgroups, so that registers can follow the work): 0-7 epilogue (warps 0-3 own edge
// positions 0-63 of the tile, warps 4-7 positions 64-127), 8 MMA issue, 9 metadata, 10-11 padding, 12-27 producers
// (8 rows per warp)
constexpr int FW_EPI_WARPS = 8, FW_MMA_WARP = 8, FW_META_WARP = 9, FW_PROD_WARP0 = 12, FW_PROD_WARPS = 16;
constexpr int FW_THREADS = (FW_PROD_WARP0 + FW_PROD_WARPS) * 32;     // 896
constexpr int FW_FLUSH_TE = 64;

// Role loops are compact on purpose (the roles of a CTA share the SM's 32 KB instruction cache).
template <int NSPLIT, bool FAST>
__global__ void __launch_bounds__(FW_THREADS, 1) gnn_edge_fwd_tc_kernel(const EdgeFwdTcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* w_img = base;                                         // NSPLIT images
    unsigned char* b_img = base + (size_t)NSPLIT * TILE_BYTES;            // [stage][split]
    TileMeta* metas = reinterpret_cast<TileMeta*>(b_img + (size_t)TC_STAGES * NSPLIT * TILE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(metas + META_STAGES);
    uint64_t* full = bars;                    // [2] producers -> MMA
    uint64_t* empty = bars + 2;               // [2] MMA -> producers
    uint64_t* tfull = bars + 4;               // [2] MMA -> epilogue
    uint64_t* tempty = bars + 6;              // [2] epilogue -> MMA
    uint64_t* mfull = bars + 8;               // [META_STAGES] meta warp -> epilogue
    uint64_t* mempty = bars + 8 + META_STAGES;   // [META_STAGES] epilogue -> meta warp
    uint64_t* wbar = bars + 8 + 2 * META_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9 + 2 * META_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_edges, TCE);
    const int nt = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);     // tiles of this CTA (>= 1)

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&full[s], FW_PROD_WARPS * 32);
            umma::mbar_init(&empty[s], 1);
            umma::mbar_init(&tfull[s], 1);
            umma::mbar_init(&tempty[s], FW_EPI_WARPS * 32);
        }
        for (int s = 0; s < META_STAGES; ++s) {
            umma::mbar_init(&mfull[s], 32);
            umma::mbar_init(&mempty[s], FW_EPI_WARPS * 32);
        }
        umma::mbar_init(wbar, 1);
        umma::fence_barrier_init();
    }
    if (warp == FW_MMA_WARP) umma::tmem_alloc(tmem_slot, 256);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < FW_EPI_WARPS) {
        umma::reg_dec<56>();
        // =========================== epilogue: thread = output channel n ===========================
        const int n = tid & 127;
        const int pos0 = (warp >> 2) * 64;      // this warp's half of the tile
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float bias = a.b2[n];
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int ms = it % META_STAGES;
            const TileMeta* M = metas + ms;
            umma::mbar_wait(&mfull[ms], (it / META_STAGES) & 1);
            umma::mbar_wait_relaxed<64>(&tfull[acc], aph);
            umma::tc_fence_after();
            float sum = 0.f;
#pragma unroll 1
            for (int cb16 = 0; cb16 < 64; cb16 += 16) {
                // sixteen positions per TMEM round trip, processed as two chunks of eight
                float v16[16];
                umma::tmem_ld16(tmem + (uint32_t)(acc * TCE) + lane_base + pos0 + cb16, v16);
                if (cb16 + 16 >= 64) {            // this thread's part of the accumulator is in registers: hand it back
                    umma::tc_fence_before();
                    umma::mbar_arrive(&tempty[acc]);
                }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int cb = cb16 + hh * 8;
                const int c0 = pos0 + cb;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = swish_tc<FAST>(v16[hh * 8 + i] + bias);
                // segmented mean over the destination-sorted positions: one pass per stored sum (segment end or
                // sub-tile end); positions past the end of the edge list carry Swish(bias) but belong to no segment
                uint32_t fm = (M->flushmask[c0 >> 5] >> (c0 & 31)) & 0xffu, todo = 0xffu;
                if (fm == 0) {                   // the common case: no sum ends inside the chunk
                    sum += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                    continue;
                }
                while (fm) {
                    const uint32_t low = fm & (0u - fm);
                    const uint32_t upto = (low << 1) - 1u;
                    const uint32_t rng = todo & upto;
                    float part = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (rng & (1u << i)) part += v[i];
                    const int pos = c0 + (31 - __clz(low));
                    M->out[pos][n] = (sum + part) * M->scale[pos];
                    sum = 0.f;
                    todo &= ~upto;
                    fm &= fm - 1;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (todo & (1u << i)) sum += v[i];
            }
            }
            umma::mbar_arrive(&mempty[ms]);
        }
    } else if (warp < FW_PROD_WARP0 && warp != FW_MMA_WARP && warp != FW_META_WARP) {
        umma::reg_dec<40>();          // padding warps of the MMA / metadata warpgroup
    } else if (warp == FW_MMA_WARP) {
        umma::reg_dec<40>();
        // =========================== MMA issue ====================================================
        if (lane == 0) load_w2_image(w_img, a.w2img, NSPLIT * TILE_BYTES, wbar);
        umma::mbar_wait(wbar, 0);
        const uint32_t idesc = umma::idesc_bf16(128, 128, 0, 0);
        const uint64_t w_d = umma::desc_sw128(umma::smem_u32(w_img), 16, 1024);
        const uint64_t b_d = umma::desc_sw128(umma::smem_u32(b_img), 16, 1024);
        constexpr uint32_t TB = TILE_BYTES >> 4;
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int s = it % TC_STAGES;
            const uint32_t ph = (it / TC_STAGES) & 1;
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            umma::mbar_wait(&tempty[acc], aph ^ 1);
            umma::mbar_wait(&full[s], ph);
            umma::tc_fence_after();
            if (umma::elect_one()) {
                const uint64_t bd = b_d + (uint64_t)((uint32_t)s * NSPLIT * TB);
                const uint32_t d = tmem + (uint32_t)(acc * TCE);
#pragma unroll
                for (int term = 0; term < (NSPLIT == 1 ? 1 : 3); ++term) {
                    const uint64_t wa = w_d + (term == 2 ? TB : 0), bb = bd + (term == 1 ? TB : 0);     // hi*hi, hi*lo, lo*hi
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t koff = (uint32_t)((k >> 2) * (128 * 128 >> 4) + (k & 3) * 2);
                        umma::mma_bf16(d, wa + (uint64_t)koff, bb + (uint64_t)koff, idesc, (term | k) ? 1u : 0u);
                    }
                }
                umma::mma_commit(&empty[s]);
                umma::mma_commit(&tfull[acc]);
            }
            __syncwarp();
        }
    } else if (warp == FW_META_WARP) {
        umma::reg_dec<40>();
        // =========================== segment metadata, META_STAGES tiles ahead =====================
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int ms = it % META_STAGES;
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            umma::mbar_wait_relaxed(&mempty[ms], ((it / META_STAGES) & 1) ^ 1);
            build_tile_meta(metas + ms, a.rowptr, a.dstv, a.srcv, a.n_edges, tile, lane, a.agg, TCH, true, a.part_head, a.part_tail, FW_FLUSH_TE);
            umma::mbar_arrive(&mfull[ms]);
        }
    } else {
        umma::reg_inc<88>();
        // =========================== producers: 8 rows per warp ====================================
        // h1[e][:] = Swish(P[dst_e] + Q[src_e]) -> bf16 (hi[/lo]) K-major swizzled image(s).  Gathers, Swish and the
        // split of a tile all happen before its stage is waited for (the converted rows sit in registers); the
        // gathers of the next tile are issued before that wait as well.
        const int pw = warp - FW_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        auto load_idx = [&](int it, int& d, int& sidx) {
            const int64_t e = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TCE + pw * 8 + lane;
            d = -1; sidx = -1;
            if (it < nt && lane < 8 && e < a.n_edges) { d = a.dstv[e]; sidx = a.srcv[e]; }
        };
        int d_cur, s_cur, d_nxt, s_nxt;
        load_idx(0, d_cur, s_cur);
        load_idx(1, d_nxt, s_nxt);
        float4 q[8], p0, p1;
        int pd0, pd1;
        auto issue_gathers = [&]() {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int sidx = __shfl_sync(0xffffffffu, s_cur, r);
                q[r] = *reinterpret_cast<const float4*>(a.pq + (int64_t)(sidx < 0 ? 0 : sidx) * (2 * TCH) + TCH + lane * 4);
            }
            // the P row of the first position and of the first position with another destination (if any)
            pd0 = __shfl_sync(0xffffffffu, d_cur, 0);
            const uint32_t chg = __ballot_sync(0xffffffffu, lane < 8 && d_cur != pd0 && d_cur >= 0);
            pd1 = chg ? __shfl_sync(0xffffffffu, d_cur, __ffs(chg) - 1) : pd0;
            p0 = *reinterpret_cast<const float4*>(a.pq + (int64_t)(pd0 < 0 ? 0 : pd0) * (2 * TCH) + lane * 4);
            p1 = *reinterpret_cast<const float4*>(a.pq + (int64_t)(pd1 < 0 ? 0 : pd1) * (2 * TCH) + lane * 4);
        };
        issue_gathers();
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int s = it % TC_STAGES;
            unsigned char* img = b_img + (size_t)s * NSPLIT * TILE_BYTES;
            uint4 hl[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int d = __shfl_sync(0xffffffffu, d_cur, r);
                if (d != pd0) {
                    if (d == pd1) p0 = p1;
                    else if (d >= 0) p0 = *reinterpret_cast<const float4*>(a.pq + (int64_t)d * (2 * TCH) + lane * 4);
                    pd0 = d;
                }
                float4 h;
                h.x = swish_tc<FAST>(p0.x + q[r].x);
                h.y = swish_tc<FAST>(p0.y + q[r].y);
                h.z = swish_tc<FAST>(p0.z + q[r].z);
                h.w = swish_tc<FAST>(p0.w + q[r].w);
                if (d < 0) h = make_float4(0.f, 0.f, 0.f, 0.f);
                if (NSPLIT == 1) {
                    hl[r].x = umma::pack_bf16(h.x, h.y);
                    hl[r].y = umma::pack_bf16(h.z, h.w);
                } else {
                    split2_bf16(h.x, h.y, hl[r].x, hl[r].z);
                    split2_bf16(h.z, h.w, hl[r].y, hl[r].w);
                }
            }
            d_cur = d_nxt; s_cur = s_nxt;
            if (it + 1 < nt) issue_gathers();
            load_idx(it + 2, d_nxt, s_nxt);
            umma::mbar_wait_relaxed<128>(&empty[s], ((it / TC_STAGES) & 1) ^ 1);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t off = lane_blk + (uint32_t)(pw * 8 + r) * 128u + ((lane_chunk ^ (uint32_t)r) << 4);
                *reinterpret_cast<uint2*>(img + off) = make_uint2(hl[r].x, hl[r].y);
                if (NSPLIT == 2) *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = make_uint2(hl[r].z, hl[r].w);
            }
            umma::fence_async_smem();
            umma::mbar_arrive(&full[s]);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == FW_MMA_WARP) umma::tmem_dealloc(tmem, 256);
}

size_t edge_fwd_tc_workspace(int64_t n_edges) {
    const int64_t tiles = ceil_div<int64_t>(n_edges > 0 ? n_edges : 1, FW_FLUSH_TE);
    return 2 * align_up((size_t)tiles * TCH * sizeof(float)) + 512;
}

// precision: 1 = bf16 hi/lo split (fp32 contract), 2 = plain bf16 + tanh.approx Swish
int launch_edge_fwd_tc(int precision, const float* pq, const int32_t* rowptr, const int32_t* dstv, const int32_t* srcv,
                       int64_t n_edges, const void* w2img, const float* b2, float* agg, void* ws_ptr, size_t ws_bytes,
                       cudaStream_t s) {
    if (n_edges <= 0) return MGB_OK;
    const int64_t tiles = ceil_div<int64_t>(n_edges, TCE);
    const int64_t subtiles = ceil_div<int64_t>(n_edges, FW_FLUSH_TE);
    Workspace ws(ws_ptr, ws_bytes);
    float* part_head = ws.take<float>((size_t)subtiles * TCH);
    float* part_tail = ws.take<float>((size_t)subtiles * TCH);
    MGB_WS_CHECK(ws);
    EdgeFwdTcArgs a{pq, rowptr, dstv, srcv, n_edges, (const unsigned char*)w2img, b2, agg, part_head, part_tail};
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    {
        ProfScope prof(PROF_EDGE_FWD, s);
        if (precision == 2) {
            constexpr size_t smem = edge_fwd_tc_smem<1>();
            MGB_CUDA(cudaFuncSetAttribute(gnn_edge_fwd_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gnn_edge_fwd_tc_kernel<1, true><<<grid, FW_THREADS, smem, s>>>(a);
        } else {
            constexpr size_t smem = edge_fwd_tc_smem<2>();
            MGB_CUDA(cudaFuncSetAttribute(gnn_edge_fwd_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gnn_edge_fwd_tc_kernel<2, false><<<grid, FW_THREADS, smem, s>>>(a);
        }
    }
    MGB_LAUNCH_CHECK();
    if (subtiles > 1) {
        segment_fixup_tc_kernel<<<(unsigned)ceil_div<int64_t>(subtiles - 1, 8), 256, 0, s>>>(rowptr, dstv, n_edges, FW_FLUSH_TE, part_head, part_tail, agg, TCH, 1);
        MGB_LAUNCH_CHECK();
    }
    return MGB_OK;
}

// ==================================================================================================
// Backward (recompute).  Per 64-edge tile, all contractions on the tensor core, transposed so that an
// epilogue thread owns one channel:
//   MMA1  D1[n][e] = sum_k W2[n][k] h1[e][k]          A = W2 image (K-major),  B = h1 tile (K-major)       N = 64
//   epi1  dz2[e][n] = dagg[dst_e][n]/deg * Swish'(D1 + b2[n])  -> DZt tile [n][e] (bf16 hi[/lo]);  db2[n] += dz2
//   MMA2  D2[k][e] = sum_n W2[n][k] dz2[e][n]         A = W2 image (MN-major), B = DZt tile (MN-major)     N = 64
//   MMA3  D3[n][k] += sum_e dz2[e][n] h1[e][k]        A = DZt tile (K-major),  B = h1 tile (MN-major)      K = 64
//         D3 = dW2 accumulates in TMEM; two D3 buffers alternate every D3_GROUP tiles and are drained into
//         the CTA's fp32 partial with round-to-nearest adds (the tensor core's own accumulation truncates,
//         which would bias a sum over thousands of K-steps past the 1e-5 contract)
//   epi2  dz1[e][k] = D2 * Swish'(P[dst_e][k] + Q[src_e][k])  -> global dz1 (by-source reduction later)
//         and the segmented SUM over the dst-sorted positions -> dP[dst]      (no atomics)
// Every buffer on the path is double-buffered (h1 tile, DZt tile, D1, D2, D3), so the five roles work on
// different tiles at the same time: producers on t+1/t+2, MMA1 on t+1, epi1 on t, MMA2/3 on t, epi2 on t-1.
// A 64-edge tile keeps all of it inside 227 KB of shared memory (W2 64 KB + h1 2x32 KB + DZt 2x32 KB in the
// hi/lo mode) and 512 TMEM columns (D1 2x64, D2 2x64, D3 2x128).
// ==================================================================================================
#ifdef MGB_TIMELINE
__device__ long long* g_timeline = nullptr;      // [role 0..4][it 0..15][event 0..3]
#define TL(role, it_, ev) do { if (blockIdx.x == 0 && (it_) < 16 && (threadIdx.x & 31) == 0 && g_timeline) g_timeline[((role) * 16 + (it_)) * 4 + (ev)] = clock64(); } while (0)
#else
#define TL(role, it_, ev) do { } while (0)
#endif

struct EdgeBwdTcArgs {
    const float* pq;
    const int32_t* rowptr;
    const int32_t* dstv;
    const int32_t* srcv;
    int64_t n_edges;
    const unsigned char* w2img;
    const float* b2;
    const float* dagg;       // [N][ld_dagg]
    int ld_dagg;
    float* dz1;              // [E][128]
    float* dpq;              // [N][256]: columns [0,128) = dP written here (pre-zeroed)
    float* part_head;
    float* part_tail;
    float* dw2_partial;      // [grid][128][128]
    float* db2_partial;      // [grid][2][128]
};

constexpr int BTE = 64;                       // edge positions per backward tile
constexpr int BW_HB = BTE * 256;              // bytes of one h1 image [64 e][128 k] bf16 (two 64-row swizzle blocks)
constexpr int BW_ZB = TCH * 128;              // bytes of one DZt image [128 n][64 e] bf16 (one 128-row swizzle block)
constexpr int BWD_NPRE = 6;                   // segments per half tile whose dagg / P rows are prefetched into shared memory
constexpr int BW_MSTAGES = 4;
// per-tile segment metadata of the backward kernel (see TileMetaT; only what the two epilogues read)
struct alignas(16) BwdMeta {
    uint32_t qoff[BTE];    // element offset of Q[src] inside pq: max(src, 0) * 256 + 128
    float* out[BTE];       // at flush positions: row (channel 0) the running sum goes to
    int segdst[BTE];       // the tile's segments in order: destination node, 1 / in-degree
    float seginv[BTE];
    uint32_t endmask[BTE / 32];
    uint32_t flushmask[BTE / 32];
    int nseg;
};
__device__ __forceinline__ void build_bwd_meta(BwdMeta* M, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ dstv,
                                               const int32_t* __restrict__ srcv, int64_t n_edges, int64_t tile, int lane,
                                               float* out_base, int ld_out, float* part_head, float* part_tail, int flush_te) {
    const int64_t e0 = tile * BTE;
    const int64_t e1 = (e0 + BTE < n_edges) ? e0 + BTE : n_edges;
    int base = 0;
#pragma unroll
    for (int j = 0; j < BTE / 32; ++j) {
        const int p = j * 32 + lane;
        const int64_t e = e0 + p;
        const bool valid = e < e1;
        const int d = valid ? dstv[e] : -1;
        const int sidx = valid ? srcv[e] : 0;
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
        M->qoff[p] = (uint32_t)sidx * (2u * TCH) + TCH;
        const int64_t f0 = e0 + (p / flush_te) * flush_te;                 // bounds of this position's sub-tile
        const int64_t f1 = (f0 + flush_te < e1) ? f0 + flush_te : e1;
        const bool inside = (int64_t)s0 >= f0 && (int64_t)s1 <= f1;
        if (is_flush)
            M->out[p] = inside ? out_base + (int64_t)d * ld_out : ((int64_t)s0 < f0 ? part_head : part_tail) + (f0 / flush_te) * TCH;
        const uint32_t em = __ballot_sync(0xffffffffu, is_end);
        const uint32_t fm = __ballot_sync(0xffffffffu, is_flush);
        const uint32_t sm = __ballot_sync(0xffffffffu, is_start);
        if (lane == 0) { M->endmask[j] = em; M->flushmask[j] = fm; }
        if (is_start) {
            const int idx = base + __popc(sm & ((1u << lane) - 1u));
            M->segdst[idx] = d;
            M->seginv[idx] = 1.0f / (float)(s1 - s0);
        }
        base += __popc(sm);
    }
    if (lane == 0) M->nseg = base;
}
// 4-byte asynchronous copy global -> shared (no register staging); visible to the issuing thread after cp_async_wait_all
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(umma::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int NSPLIT>
constexpr size_t edge_bwd_tc_smem() {
    return 1024 + (size_t)NSPLIT * TILE_BYTES + (size_t)2 * NSPLIT * BW_HB + (size_t)2 * NSPLIT * BW_ZB +
           BW_MSTAGES * sizeof(BwdMeta) + 8 * BWD_NPRE * TCH * sizeof(float) + 512;
}

// warp roles of the backward kernel: 0-3 epilogue 1 (dz2, dW2 drain); 4-11 epilogue 2 (dz1, dP; warps 4-7 own edge
// positions 0-31 of the tile, warps 8-11 positions 32-63); 12 MMA; 13 metadata; 14-21 producers (two groups of four
// warps, group g fills stage g for the tiles with (it & 1) == g, 16 rows per warp)
constexpr int BW_E1_WARPS = 8, BW_E2_WARPS = 8;
constexpr int BW_MMA_WARP = 16, BW_META_WARP = 17, BW_PROD_WARP0 = 20, BW_PROD_WARPS = 8;     // warps 18, 19 idle (warpgroup padding)
constexpr int BW_THREADS = (BW_PROD_WARP0 + BW_PROD_WARPS) * 32;     // 704
constexpr int BW_FLUSH_TE = 32;
constexpr int BW_D3_GROUP = 8;

template <int NSPLIT, bool FAST>
__global__ void __launch_bounds__(BW_THREADS, 1) gnn_edge_bwd_tc_kernel(const EdgeBwdTcArgs a) {
    constexpr int D3_GROUP = FAST ? (1 << 30) : BW_D3_GROUP;
    constexpr int NTERM = NSPLIT == 1 ? 1 : 3;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* w_img = base;                                              // [split]
    unsigned char* h_img = w_img + (size_t)NSPLIT * TILE_BYTES;               // [stage][split]   h1[e][k]
    unsigned char* dz_img = h_img + (size_t)2 * NSPLIT * BW_HB;               // [stage][split]   DZt[n][e]
    BwdMeta* metas = reinterpret_cast<BwdMeta*>(dz_img + (size_t)2 * NSPLIT * BW_ZB);
    float* gtab_all = reinterpret_cast<float*>(metas + BW_MSTAGES);    // [2 buffers][2 halves][BWD_NPRE][128]  dagg[segdst]
    float* ptab_all = gtab_all + 4 * BWD_NPRE * TCH;                   // [2 buffers][2 halves][BWD_NPRE][128]  P[segdst]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ptab_all + 4 * BWD_NPRE * TCH);
    uint64_t* h_full = bars;             // [2] producers -> MMA (h1 tile ready)
    uint64_t* h_empty = bars + 2;        // [2] MMA3 done -> producers
    uint64_t* d1_full = bars + 4;        // [2] MMA1 done -> epilogue 1
    uint64_t* d1_empty = bars + 6;       // [2] epilogue 1 drained D1 -> MMA
    uint64_t* dz_full = bars + 8;        // [2] epilogue 1 wrote DZt -> MMA
    uint64_t* dz_empty = bars + 10;      // [2] MMA2+MMA3 done reading DZt -> epilogue 1
    uint64_t* d2_full = bars + 12;       // [2] MMA2 done -> epilogue 2
    uint64_t* d2_empty = bars + 14;      // [2] epilogue 2 drained D2 -> MMA
    uint64_t* d3_full = bars + 16;       // [2] a D3 group is complete -> epilogue 1
    uint64_t* d3_empty = bars + 18;      // [2] epilogue 1 drained D3 buffer -> MMA
    uint64_t* mfull = bars + 20;         // [BW_MSTAGES]
    uint64_t* mempty = bars + 20 + BW_MSTAGES;
    uint64_t* wbar = bars + 20 + 2 * BW_MSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21 + 2 * BW_MSTAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_edges, BTE);
    const int nt = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);     // tiles of this CTA (>= 1)

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&h_full[s], BW_PROD_WARPS * 32);
            umma::mbar_init(&h_empty[s], 1);
            umma::mbar_init(&d1_full[s], 1);
            umma::mbar_init(&d1_empty[s], BW_E1_WARPS * 32);
            umma::mbar_init(&dz_full[s], BW_E1_WARPS * 32);
            umma::mbar_init(&dz_empty[s], 1);
            umma::mbar_init(&d2_full[s], 1);
            umma::mbar_init(&d2_empty[s], BW_E2_WARPS * 32);
            umma::mbar_init(&d3_full[s], 1);
            umma::mbar_init(&d3_empty[s], BW_E1_WARPS * 32);
        }
        for (int s = 0; s < BW_MSTAGES; ++s) {
            umma::mbar_init(&mfull[s], 32);
            umma::mbar_init(&mempty[s], (BW_E1_WARPS + BW_E2_WARPS) * 32);
        }
        umma::mbar_init(wbar, 1);
        umma::fence_barrier_init();
    }
    if (warp == BW_MMA_WARP) umma::tmem_alloc(tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_d1 = tmem, tm_d2 = tmem + 128, tm_d3 = tmem + 256;   // D1[s] at +64 s, D2[s] at +128 + 64 s, D3[buf] at +256 + 128 buf

    // Registers follow the work: the producers hold a whole tile of gathered rows plus its converted image, the MMA and
    // metadata warps need next to nothing (setmaxnreg moves registers between whole warpgroups; 896 x 72 in total).
    if (warp < BW_E1_WARPS) {
        umma::reg_dec<56>();
        // =========================== epilogue 1: thread = channel n; warps 0-3 positions 0-31, warps 4-7 positions 32-63
        const int half = warp >> 2;
        const int pos0 = half * 32;
        const int n = tid & 127;
        const float bias = a.b2[n];
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float* gtab = gtab_all + half * BWD_NPRE * TCH + n;    // [buffer][half][segment][channel]: own column per thread
        float* dw_out = a.dw2_partial + ((int64_t)blockIdx.x * TCH + n) * TCH + half * 64;
        const float* dagg_n = a.dagg + n;
        float db = 0.f;
        auto drain_d3 = [&](int grp) {
            // a finished D3 group -> this CTA's fp32 partial (round-to-nearest adds); this half owns 64 of the 128 columns
            const int buf = grp & 1;
            umma::mbar_wait(&d3_full[buf], (grp >> 1) & 1);
            umma::tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 16) {
                float4* o = reinterpret_cast<float4*>(dw_out + c0);
                // the running partial is requested before the accumulator columns: its L2 latency overlaps the TMEM round trip
                // (ncu: 5.6 % of the kernel's stall samples sat on these loads behind the tcgen05.ld)
                float4 old[4];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) old[q4] = grp > 0 ? o[q4] : make_float4(0.f, 0.f, 0.f, 0.f);
                float v[16];
                umma::tmem_ld16(tm_d3 + (uint32_t)(buf * 128 + half * 64) + lane_base + c0, v);
                if (c0 + 16 >= 64) {
                    umma::tc_fence_before();
                    umma::mbar_arrive(&d3_empty[buf]);
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    float4 w = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
                    w.x += old[q4].x; w.y += old[q4].y; w.z += old[q4].z; w.w += old[q4].w;
                    o[q4] = w;
                }
            }
        };
        // dagg[dst] of the first BWD_NPRE segments of this half goes global -> shared (cp.async, no registers) one tile
        // ahead, into the other buffer of the table
        auto prefetch_tab = [&](const BwdMeta* M, int buf) {
            const int nseg = M->nseg;
            const int j0 = half ? __popc(M->endmask[0]) : 0;
            float* dstp = gtab + buf * (2 * BWD_NPRE * TCH);
#pragma unroll
            for (int j = 0; j < BWD_NPRE; ++j) {
                int jj = j0 + j;
                jj = jj < nseg ? jj : nseg - 1;
                cp_async4(dstp + j * TCH, dagg_n + (int64_t)M->segdst[jj] * a.ld_dagg);
            }
        };
        umma::mbar_wait(&mfull[0], 0);
        prefetch_tab(metas, 0);
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int ms = it % BW_MSTAGES;
            const BwdMeta* M = metas + ms;
            unsigned char* dz_row = dz_img + (size_t)s * NSPLIT * BW_ZB + n * 128;
            const float* gt = gtab + s * (2 * BWD_NPRE * TCH);
            const int nseg = M->nseg;
            const int j0 = half ? __popc(M->endmask[0]) : 0;     // segment that contains position pos0
            cp_async_wait_all();
            if (it + 1 < nt) {
                const int ms1 = (it + 1) % BW_MSTAGES;
                umma::mbar_wait(&mfull[ms1], ((it + 1) / BW_MSTAGES) & 1);
                prefetch_tab(metas + ms1, s ^ 1);
            }
            if (warp == 0) TL(2, it, 0);
            umma::mbar_wait(&d1_full[s], ph);
            if (warp == 0) TL(2, it, 1);
            umma::mbar_wait(&dz_empty[s], ph ^ 1);
            umma::tc_fence_after();
            if (warp == 0) TL(2, it, 2);
            int j = j0;
            float g = j < nseg ? gt[0] * M->seginv[j] : 0.f;
            const uint32_t emw = M->endmask[half];
#pragma unroll 1
            for (int cb16 = 0; cb16 < 32; cb16 += 16) {
                // sixteen positions per TMEM round trip, processed as two chunks of eight
                float v16[16];
                umma::tmem_ld16(tm_d1 + (uint32_t)(s * BTE) + lane_base + pos0 + cb16, v16);
                if (cb16 + 16 >= 32) {           // this thread's part of D1[s] is in registers: MMA1 of tile it+2 may overwrite it
                    umma::tc_fence_before();
                    umma::mbar_arrive(&d1_empty[s]);
                }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int cb = cb16 + hh * 8;
                const int c0 = pos0 + cb;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = swish_grad_tc<FAST>(v16[hh * 8 + i] + bias);
                // dagg[dst]/deg is constant along a segment: one pass per segment that intersects the chunk (usually one)
                uint32_t em = (emw >> cb) & 0xffu, todo = 0xffu;
                if (em == 0) {                   // the common case: one segment covers the chunk
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] *= g;
                } else
                while (true) {
                    const uint32_t upto = em ? (((em & (0u - em)) << 1) - 1u) : 0xffu;
                    const uint32_t rng = todo & upto;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (rng & (1u << i)) v[i] *= g;
                    if (!em) break;
                    todo &= ~upto;
                    em &= em - 1;
                    ++j;
                    g = j >= nseg ? 0.f : (j - j0 < BWD_NPRE ? gt[(j - j0) * TCH] : dagg_n[(int64_t)M->segdst[j] * a.ld_dagg]) * M->seginv[j < nseg ? j : 0];
                }
                db += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                const uint32_t off = (uint32_t)(((c0 >> 3) ^ (n & 7)) << 4);
                uint4 hi, lo;
                if (NSPLIT == 1) {
                    hi.x = umma::pack_bf16(v[0], v[1]);
                    hi.y = umma::pack_bf16(v[2], v[3]);
                    hi.z = umma::pack_bf16(v[4], v[5]);
                    hi.w = umma::pack_bf16(v[6], v[7]);
                    *reinterpret_cast<uint4*>(dz_row + off) = hi;
                } else {
                    split2_bf16(v[0], v[1], hi.x, lo.x);
                    split2_bf16(v[2], v[3], hi.y, lo.y);
                    split2_bf16(v[4], v[5], hi.z, lo.z);
                    split2_bf16(v[6], v[7], hi.w, lo.w);
                    *reinterpret_cast<uint4*>(dz_row + off) = hi;
                    *reinterpret_cast<uint4*>(dz_row + BW_ZB + off) = lo;
                }
            }
            }
            umma::fence_async_smem();
            umma::tc_fence_before();
            umma::mbar_arrive(&dz_full[s]);
            if (warp == 0) TL(2, it, 3);
            umma::mbar_arrive(&mempty[ms]);
            // D3 drains run one tile late (MMA3 of the group's last tile has finished by then), except at the very end
            if (it > 0 && ((it - 1) % D3_GROUP) == D3_GROUP - 1) drain_d3((it - 1) / D3_GROUP);
            if (it == nt - 1) drain_d3(it / D3_GROUP);
        }
        a.db2_partial[((int64_t)blockIdx.x * 2 + half) * TCH + n] = db;
    } else if (warp < BW_E1_WARPS + BW_E2_WARPS) {
        // =========================== epilogue 2: thread = channel k; warps 8-11 positions 0-31, warps 12-15 positions 32-63
        const int half = (warp - BW_E1_WARPS) >> 2;
        const int pos0 = half * 32;
        const int n = tid & 127;
        float* ptab = ptab_all + half * BWD_NPRE * TCH + n;     // [buffer][half][segment][channel]: own column per thread
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float* pq_n = a.pq + n;
        // P[dst] of the first BWD_NPRE segments of this half: global -> shared (cp.async) one tile ahead
        auto prefetch_tab = [&](const BwdMeta* M, int buf) {
            const int nseg = M->nseg;
            const int j0 = half ? __popc(M->endmask[0]) : 0;
            float* dstp = ptab + buf * (2 * BWD_NPRE * TCH);
#pragma unroll
            for (int j = 0; j < BWD_NPRE; ++j) {
                int jj = j0 + j;
                jj = jj < nseg ? jj : nseg - 1;
                cp_async4(dstp + j * TCH, pq_n + (int64_t)M->segdst[jj] * (2 * TCH));
            }
        };
        umma::mbar_wait(&mfull[0], 0);
        prefetch_tab(metas, 0);
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int ms = it % BW_MSTAGES;
            const BwdMeta* M = metas + ms;
            const float* pt = ptab + s * (2 * BWD_NPRE * TCH);
            const int nseg = M->nseg;
            const int j0 = half ? __popc(M->endmask[0]) : 0;     // segment that contains position pos0
            // Q re-gather, two 8-position chunks ahead of the chunk being processed; the first two are issued before
            // MMA2 of the tile is waited for
            float qn1[8], qn2[8];
            {
                const uint4 o0 = *reinterpret_cast<const uint4*>(&M->qoff[pos0]), o1 = *reinterpret_cast<const uint4*>(&M->qoff[pos0 + 4]);
                const uint4 o2 = *reinterpret_cast<const uint4*>(&M->qoff[pos0 + 8]), o3 = *reinterpret_cast<const uint4*>(&M->qoff[pos0 + 12]);
                qn1[0] = pq_n[o0.x]; qn1[1] = pq_n[o0.y]; qn1[2] = pq_n[o0.z]; qn1[3] = pq_n[o0.w];
                qn1[4] = pq_n[o1.x]; qn1[5] = pq_n[o1.y]; qn1[6] = pq_n[o1.z]; qn1[7] = pq_n[o1.w];
                qn2[0] = pq_n[o2.x]; qn2[1] = pq_n[o2.y]; qn2[2] = pq_n[o2.z]; qn2[3] = pq_n[o2.w];
                qn2[4] = pq_n[o3.x]; qn2[5] = pq_n[o3.y]; qn2[6] = pq_n[o3.z]; qn2[7] = pq_n[o3.w];
            }
            cp_async_wait_all();
            if (it + 1 < nt) {
                const int ms1 = (it + 1) % BW_MSTAGES;
                umma::mbar_wait(&mfull[ms1], ((it + 1) / BW_MSTAGES) & 1);
                prefetch_tab(metas + ms1, s ^ 1);
            }
            int j = j0;
            float pk = j < nseg ? pt[0] : 0.f;
            float sum = 0.f;
            const uint32_t emw = M->endmask[half], fmw = M->flushmask[half];
            float* dz_tile = a.dz1 + (tile * BTE + pos0) * TCH + n;     // dz1 is padded to whole tiles: no bounds checks
            if ((warp & 3) == 0) TL(3 + half, it, 0);
            umma::mbar_wait_relaxed<64>(&d2_full[s], ph);
            umma::tc_fence_after();
            if ((warp & 3) == 0) TL(3 + half, it, 1);
#pragma unroll 1
            for (int cb = 0; cb < 32; cb += 8) {
                float q[8], v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { q[i] = qn1[i]; qn1[i] = qn2[i]; }
                if (cb + 16 < 32) {
                    const uint4 o0 = *reinterpret_cast<const uint4*>(&M->qoff[pos0 + cb + 16]);
                    const uint4 o1 = *reinterpret_cast<const uint4*>(&M->qoff[pos0 + cb + 20]);
                    qn2[0] = pq_n[o0.x]; qn2[1] = pq_n[o0.y]; qn2[2] = pq_n[o0.z]; qn2[3] = pq_n[o0.w];
                    qn2[4] = pq_n[o1.x]; qn2[5] = pq_n[o1.y]; qn2[6] = pq_n[o1.z]; qn2[7] = pq_n[o1.w];
                }
                umma::tmem_ld8(tm_d2 + (uint32_t)(s * BTE) + lane_base + pos0 + cb, v);
                if (cb + 8 >= 32) {
                    umma::tc_fence_before();
                    umma::mbar_arrive(&d2_empty[s]);
                }
                // z1 = P[dst] + Q[src]; P is constant along a segment: one pass per segment that intersects the chunk.
                // Positions past the end of the edge list carry D2 = 0 and belong to no segment: they store zeros
                // into the padding of dz1 and add nothing to any sum.
                uint32_t em = (emw >> cb) & 0xffu, todo = 0xffu;
                if (em == 0) {                   // the common case: one segment covers the chunk
#pragma unroll
                    for (int i = 0; i < 8; ++i) q[i] += pk;
                } else
                while (true) {
                    const uint32_t upto = em ? (((em & (0u - em)) << 1) - 1u) : 0xffu;
                    const uint32_t rng = todo & upto;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (rng & (1u << i)) q[i] += pk;
                    if (!em) break;
                    todo &= ~upto;
                    em &= em - 1;
                    ++j;
                    pk = j >= nseg ? 0.f : (j - j0 < BWD_NPRE ? pt[(j - j0) * TCH] : pq_n[(int64_t)M->segdst[j] * (2 * TCH)]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= swish_grad_tc<FAST>(q[i]);
                float* dzrow = dz_tile + cb * TCH;
#pragma unroll
                for (int i = 0; i < 8; ++i) dzrow[i * TCH] = v[i];
                // segmented sum over the destination-sorted positions: one pass per stored sum (segment end or sub-tile end)
                uint32_t fm = (fmw >> cb) & 0xffu;
                todo = 0xffu;
                if (fm == 0) {                   // the common case: no sum ends inside the chunk
                    sum += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                    continue;
                }
                while (fm) {
                    const uint32_t low = fm & (0u - fm);
                    const uint32_t upto = (low << 1) - 1u;
                    const uint32_t rng = todo & upto;
                    float part = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (rng & (1u << i)) part += v[i];
                    M->out[pos0 + cb + (31 - __clz(low))][n] = sum + part;
                    sum = 0.f;
                    todo &= ~upto;
                    fm &= fm - 1;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (todo & (1u << i)) sum += v[i];
            }
            if ((warp & 3) == 0) TL(3 + half, it, 2);
            umma::mbar_arrive(&mempty[ms]);
        }
    } else if (warp < BW_PROD_WARP0 && warp != BW_MMA_WARP && warp != BW_META_WARP) {
        umma::reg_dec<40>();          // padding warps of the MMA / metadata warpgroup
    } else if (warp == BW_MMA_WARP) {
        umma::reg_dec<40>();
        // =========================== MMA issue ====================================================
        if (lane == 0) load_w2_image(w_img, a.w2img, NSPLIT * TILE_BYTES, wbar);
        umma::mbar_wait(wbar, 0);
        const uint32_t id_1 = umma::idesc_bf16(128, BTE, 0, 0);     // A K-major,  B K-major   (MMA1)
        const uint32_t id_2 = umma::idesc_bf16(128, BTE, 1, 1);     // A MN-major, B MN-major  (MMA2)
        const uint32_t id_3 = umma::idesc_bf16(128, 128, 0, 1);     // A K-major,  B MN-major  (MMA3)
        // descriptors of k-step 0; a k-step only moves the 14-bit start-address field (bytes >> 4), so the
        // per-MMA work is one add per operand
        const uint64_t w_k = umma::desc_sw128(umma::smem_u32(w_img), 16, 1024);             // W2, K-major
        const uint64_t w_m = umma::desc_sw128(umma::smem_u32(w_img), 128 * 128, 1024);      // W2, MN-major
        const uint64_t h_k = umma::desc_sw128(umma::smem_u32(h_img), 16, 1024);             // h1[e][k], K-major  (K = k)
        const uint64_t h_m = umma::desc_sw128(umma::smem_u32(h_img), BTE * 128, 1024);      // h1[e][k], MN-major (K = e)
        const uint64_t z_k = umma::desc_sw128(umma::smem_u32(dz_img), 16, 1024);            // DZt[n][e], K-major (K = e)
        const uint64_t z_m = umma::desc_sw128(umma::smem_u32(dz_img), 128 * 128, 1024);     // DZt[n][e], MN-major (K = n)
        constexpr uint32_t WT = TILE_BYTES >> 4, HS = (NSPLIT * BW_HB) >> 4, HT = BW_HB >> 4, ZS = (NSPLIT * BW_ZB) >> 4,
                           ZT = BW_ZB >> 4;
        // MMA1 of tile it+1 is issued before MMA2/MMA3 of tile it: the tensor pipe works on the next tile while
        // epilogue 1 runs on this one
#pragma unroll 1
        for (int it = -1; it < nt; ++it) {
            if (it + 1 < nt) {
                const int it1 = it + 1, s1 = it1 & 1;
                umma::mbar_wait(&h_full[s1], (it1 >> 1) & 1);
                umma::mbar_wait(&d1_empty[s1], ((it1 >> 1) & 1) ^ 1);
                umma::tc_fence_after();
                TL(1, it1, 0);
                if (umma::elect_one()) {
                    const uint64_t hb = h_k + (uint64_t)(s1 * HS);
                    const uint32_t d = tm_d1 + (uint32_t)(s1 * BTE);
#pragma unroll
                    for (int term = 0; term < NTERM; ++term) {
                        const uint64_t wa = w_k + (term == 2 ? WT : 0), hh = hb + (term == 1 ? HT : 0);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma::mma_bf16(d, wa + (uint64_t)((k >> 2) * (128 * 128 >> 4) + (k & 3) * 2),
                                           hh + (uint64_t)((k >> 2) * (BTE * 128 >> 4) + (k & 3) * 2), id_1, (term | k) ? 1u : 0u);
                    }
                    umma::mma_commit(&d1_full[s1]);
                }
                __syncwarp();
            }
            if (it < 0) continue;
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int grp = it / D3_GROUP, buf = grp & 1;
            const bool first_in_group = (it % D3_GROUP) == 0;
            const bool last = it == nt - 1;
            umma::mbar_wait(&dz_full[s], ph);
            umma::mbar_wait(&d2_empty[s], ph ^ 1);
            if (first_in_group) umma::mbar_wait(&d3_empty[buf], ((grp >> 1) & 1) ^ 1);
            umma::tc_fence_after();
            TL(1, it, 1);
            if (umma::elect_one()) {
                const uint64_t hb = h_m + (uint64_t)(s * HS), zk = z_k + (uint64_t)(s * ZS), zm = z_m + (uint64_t)(s * ZS);
                const uint32_t d2 = tm_d2 + (uint32_t)(s * BTE), d3 = tm_d3 + (uint32_t)(buf * 128);
#pragma unroll
                for (int term = 0; term < NTERM; ++term) {
                    const uint64_t wa = w_m + (term == 2 ? WT : 0), zz = zm + (term == 1 ? ZT : 0);
#pragma unroll
                    for (int k = 0; k < 8; ++k)     // K = n: 16 rows = 2048 bytes per step
                        umma::mma_bf16(d2, wa + (uint64_t)(k * 128), zz + (uint64_t)(k * 128), id_2, (term | k) ? 1u : 0u);
                }
                umma::mma_commit(&d2_full[s]);
#pragma unroll
                for (int term = 0; term < NTERM; ++term) {
                    const uint64_t za = zk + (term == 2 ? ZT : 0), hh = hb + (term == 1 ? HT : 0);
#pragma unroll
                    for (int k = 0; k < BTE / 16; ++k)     // K = e
                        umma::mma_bf16(d3, za + (uint64_t)(k * 2), hh + (uint64_t)(k * 128), id_3, (term | k) ? 1u : (first_in_group ? 0u : 1u));
                }
                umma::mma_commit(&h_empty[s]);
                umma::mma_commit(&dz_empty[s]);
                if ((it % D3_GROUP) == D3_GROUP - 1 || last) umma::mma_commit(&d3_full[buf]);
            }
            __syncwarp();
            TL(1, it, 2);
        }
    } else if (warp == BW_META_WARP) {
        umma::reg_dec<40>();
        for (int it = 0; it < nt; ++it) {
            const int ms = it % BW_MSTAGES;
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            umma::mbar_wait_relaxed(&mempty[ms], ((it / BW_MSTAGES) & 1) ^ 1);
            build_bwd_meta(metas + ms, a.rowptr, a.dstv, a.srcv, a.n_edges, tile, lane, a.dpq, 2 * TCH, a.part_head, a.part_tail, BW_FLUSH_TE);
            umma::mbar_arrive(&mfull[ms]);
        }
    } else {
        umma::reg_inc<104>();
        // =========================== producers: 8 rows per warp ====================================
        // h1[e][:] = Swish(P[dst_e] + Q[src_e]) -> bf16 (hi[/lo]) K-major swizzled image(s).  The eight Q-row gathers of a
        // tile (one 512-byte coalesced row per load instruction) and its edge indices are issued while the previous
        // tiles are still in the pipeline, so that only Swish + split + stores follow the stage's release.
        const int pw = warp - BW_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * ((uint32_t)BTE * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        auto load_idx = [&](int it, int& d, int& sidx) {
            const int64_t e = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * BTE + pw * 8 + lane;
            d = -1; sidx = -1;
            if (it < nt && lane < 8 && e < a.n_edges) { d = a.dstv[e]; sidx = a.srcv[e]; }
        };
        int d_cur, s_cur, d_nxt, s_nxt;
        load_idx(0, d_cur, s_cur);
        load_idx(1, d_nxt, s_nxt);
        float4 q[8], p0, p1;
        int pd0, pd1;
        auto issue_gathers = [&]() {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int sidx = __shfl_sync(0xffffffffu, s_cur, r);
                q[r] = *reinterpret_cast<const float4*>(a.pq + (int64_t)(sidx < 0 ? 0 : sidx) * (2 * TCH) + TCH + lane * 4);
            }
            // the P row of the first position and of the first position with another destination (if any)
            pd0 = __shfl_sync(0xffffffffu, d_cur, 0);
            const uint32_t chg = __ballot_sync(0xffffffffu, lane < 8 && d_cur != pd0 && d_cur >= 0);
            pd1 = chg ? __shfl_sync(0xffffffffu, d_cur, __ffs(chg) - 1) : pd0;
            p0 = *reinterpret_cast<const float4*>(a.pq + (int64_t)(pd0 < 0 ? 0 : pd0) * (2 * TCH) + lane * 4);
            p1 = *reinterpret_cast<const float4*>(a.pq + (int64_t)(pd1 < 0 ? 0 : pd1) * (2 * TCH) + lane * 4);
        };
        issue_gathers();
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int s = it & 1;
            unsigned char* img = h_img + (size_t)s * NSPLIT * BW_HB;
            if (pw == 0) TL(0, it, 0);
            // Swish + split run before the stage is waited for (results stay in the registers the gathered rows came
            // in), so only the stores follow the release of the stage
            uint4 hl[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int d = __shfl_sync(0xffffffffu, d_cur, r);
                if (d != pd0) {
                    if (d == pd1) p0 = p1;
                    else if (d >= 0) p0 = *reinterpret_cast<const float4*>(a.pq + (int64_t)d * (2 * TCH) + lane * 4);
                    pd0 = d;
                }
                float4 h;
                h.x = swish_tc<FAST>(p0.x + q[r].x);
                h.y = swish_tc<FAST>(p0.y + q[r].y);
                h.z = swish_tc<FAST>(p0.z + q[r].z);
                h.w = swish_tc<FAST>(p0.w + q[r].w);
                if (d < 0) h = make_float4(0.f, 0.f, 0.f, 0.f);
                if (NSPLIT == 1) {
                    hl[r].x = umma::pack_bf16(h.x, h.y);
                    hl[r].y = umma::pack_bf16(h.z, h.w);
                } else {
                    split2_bf16(h.x, h.y, hl[r].x, hl[r].z);
                    split2_bf16(h.z, h.w, hl[r].y, hl[r].w);
                }
            }
            // the gathers of the next tile go out before this tile's stage is waited for: their latency overlaps the wait
            d_cur = d_nxt; s_cur = s_nxt;
            if (it + 1 < nt) issue_gathers();
            load_idx(it + 2, d_nxt, s_nxt);
            umma::mbar_wait_relaxed<128>(&h_empty[s], ((it >> 1) & 1) ^ 1);
            if (pw == 0) TL(0, it, 2);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t off = lane_blk + (uint32_t)(pw * 8 + r) * 128u + ((lane_chunk ^ (uint32_t)r) << 4);
                *reinterpret_cast<uint2*>(img + off) = make_uint2(hl[r].x, hl[r].y);
                if (NSPLIT == 2) *reinterpret_cast<uint2*>(img + BW_HB + off) = make_uint2(hl[r].z, hl[r].w);
            }
            umma::fence_async_smem();
            umma::mbar_arrive(&h_full[s]);
            if (pw == 0) TL(0, it, 1);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == BW_MMA_WARP) umma::tmem_dealloc(tmem, 512);
}

#ifdef MGB_TIMELINE
int set_timeline_buffer(long long* p) {
    MGB_CUDA(cudaMemcpyToSymbol(g_timeline, &p, sizeof(p)));
    return MGB_OK;
}
#endif

int edge_bwd_tc_grid(int64_t n_edges) {
    const int64_t tiles = ceil_div<int64_t>(n_edges > 0 ? n_edges : 1, BTE);
    return (int)(tiles < sm_count() ? tiles : sm_count());
}

size_t edge_bwd_tc_workspace(int64_t n_edges) {
    const int64_t tiles = ceil_div<int64_t>(n_edges > 0 ? n_edges : 1, BW_FLUSH_TE);
    const int grid = edge_bwd_tc_grid(n_edges);
    return 2 * align_up((size_t)tiles * TCH * sizeof(float)) + align_up((size_t)grid * TCH * TCH * sizeof(float)) +
           align_up((size_t)grid * 2 * TCH * sizeof(float)) + 1024;
}

// out[i] (+)= sum over the CTA partials.  Block = 32 outputs x 8 groups; group g sums the partials p = g, g+8, ... with four
// independent loads in flight, then a fixed-order shared-memory reduction over the groups (deterministic).
__global__ void __launch_bounds__(256)
sum_partials_tc_kernel(const float* __restrict__ partial, int n_parts, int64_t count, float* __restrict__ out, int accumulate) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < count) {
        int p = g;
        for (; p + 24 < n_parts; p += 32) {
            const float v0 = partial[(int64_t)p * count + i], v1 = partial[(int64_t)(p + 8) * count + i];
            const float v2 = partial[(int64_t)(p + 16) * count + i], v3 = partial[(int64_t)(p + 24) * count + i];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; p < n_parts; p += 8) s += partial[(int64_t)p * count + i];
    }
    red[g][lane] = s;
    __syncthreads();
    if (g == 0 && i < count) {
        float t = red[0][lane];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += red[q][lane];
        out[i] = accumulate ? out[i] + t : t;
    }
}

int launch_edge_bwd_tc(int precision, const float* pq, const int32_t* rowptr, const int32_t* dstv, const int32_t* srcv,
                       int64_t n_edges, const void* w2img, const float* b2, const float* dagg, int ld_dagg, float* dz1,
                       float* dpq, float* dW2, float* db2, int accumulate, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    if (n_edges <= 0) return MGB_OK;
    const int64_t subtiles = ceil_div<int64_t>(n_edges, BW_FLUSH_TE);
    const int grid = edge_bwd_tc_grid(n_edges);
    Workspace ws(ws_ptr, ws_bytes);
    float* part_head = ws.take<float>((size_t)subtiles * TCH);
    float* part_tail = ws.take<float>((size_t)subtiles * TCH);
    float* dw2_part = ws.take<float>((size_t)grid * TCH * TCH);
    float* db2_part = ws.take<float>((size_t)grid * 2 * TCH);
    MGB_WS_CHECK(ws);
    EdgeBwdTcArgs a{pq, rowptr, dstv, srcv, n_edges, (const unsigned char*)w2img, b2, dagg, ld_dagg, dz1, dpq,
                    part_head, part_tail, dw2_part, db2_part};
    {
        ProfScope prof(PROF_EDGE_BWD, s);
        if (precision == 2) {
            constexpr size_t smem = edge_bwd_tc_smem<1>();
            MGB_CUDA(cudaFuncSetAttribute(gnn_edge_bwd_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gnn_edge_bwd_tc_kernel<1, true><<<grid, BW_THREADS, smem, s>>>(a);
        } else {
            constexpr size_t smem = edge_bwd_tc_smem<2>();
            MGB_CUDA(cudaFuncSetAttribute(gnn_edge_bwd_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gnn_edge_bwd_tc_kernel<2, false><<<grid, BW_THREADS, smem, s>>>(a);
        }
    }
    MGB_LAUNCH_CHECK();
    if (subtiles > 1) {
        segment_fixup_tc_kernel<<<(unsigned)ceil_div<int64_t>(subtiles - 1, 8), 256, 0, s>>>(rowptr, dstv, n_edges, BW_FLUSH_TE, part_head, part_tail, dpq, 2 * TCH, 0);
        MGB_LAUNCH_CHECK();
    }
    sum_partials_tc_kernel<<<TCH * TCH / 32, 256, 0, s>>>(dw2_part, grid, (int64_t)TCH * TCH, dW2, accumulate);
    MGB_LAUNCH_CHECK();
    sum_partials_tc_kernel<<<TCH / 32, 256, 0, s>>>(db2_part, 2 * grid, TCH, db2, accumulate);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
