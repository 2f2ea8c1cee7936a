// magnet_b200 — fused edge kernel of MAgNet's InteractionNetwork on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Reference: InteractionNetwork.message + aggregate (models/magnet_gnn.py:70-90, edge_fn :61-68, aggr='mean' :54):
//     m_e   = LayerNorm(MLP5(cat[x_i, x_j, e_e]))        384 -> 128 -> 128 -> 128 -> 128 -> 128, ReLU between
//     agg_i = mean_{e -> i} m_e                           at i = edge_index[1] (UNSORTED in MAgNetGNN graphs)
// The first Linear is factorised: W0 [x_i, x_j, e] = P[i] + Q[j] + We e with P|Q = x [W0_i | W0_j]^T + [b0 | 0] computed per
// NODE (linear_tc.cu); what remains per EDGE is a chain of five 128x128 contractions, which never leaves the SM here:
//   producers   gather e_features rows through the aggregation plan (position -> COO edge id), scale by 2^l (the
//               reference doubles e_features every layer and never updates them), split into fp16 hi | lo and
//               write a K-major operand image (as mlp_chain_tc.cu)
//               and initialise the accumulator with P[dst_e][n] + Q[src_e][n] (tcgen05.st, thread = channel n)
//   layer 0     D^T[n][e] += sum_k We[n][k] e[e][k];  epilogue (thread = channel n): ReLU, next operand written IN PLACE
//               as an MN-major image
//   layers 1-3  as mlp_chain_tc.cu (bias, ReLU, MN-major image)
//   layer 4     y = D + b4;  LayerNorm over the 128 channels of an edge = across the 128 epilogue threads: y goes through
//               the (now free) operand tile as an fp32 [e][n] staging buffer, two threads per edge compute mean / rstd
//               exactly (two-pass, values in registers), then every thread normalises its channel from TMEM and runs the
//               segmented mean over the destination-sorted positions in registers (segmeta.cuh; no atomics, fixed order)
// Two 128-edge tiles ping-pong between the MMA warp and the epilogue warps; one layer's weight images (64 KB) are
// re-loaded from L2 per layer with one bulk async copy.  HBM sees e_features once (512 B per edge) and agg once.
// NSPLIT = 2: operands split into two fp16 values, three MMA terms (1e-5 contract); NSPLIT = 1: plain bf16 (1e-2).
#include "internal.cuh"
#include "tc_common.cuh"
#include "segmeta.cuh"

namespace mgb {

constexpr int IE_TE = 128;                       // edge positions per tile (MMA N)
constexpr int IE_L = 5;                          // Linear layers of edge_fn
constexpr int IE_EPI_WARPS = 8, IE_PROD_WARPS = 8;
constexpr int IE_MMA_WARP = IE_EPI_WARPS, IE_META_WARP = IE_EPI_WARPS + 1, IE_PROD_WARP0 = IE_EPI_WARPS + 2;
constexpr int IE_THREADS = (IE_PROD_WARP0 + IE_PROD_WARPS) * 32;      // 576
constexpr int IE_FLUSH = 64;                     // positions per epilogue warp = granularity of the stored partial sums
constexpr int IE_MSLOTS = 4;                     // metadata slots: (pair parity, tile of the pair)
using IeMeta = TileMetaT<IE_TE>;
constexpr size_t IN_EDGE_SMEM = 1024 + (size_t)2 * TILE_BYTES + (size_t)4 * TILE_BYTES + IE_MSLOTS * sizeof(IeMeta) +
                                4 * IE_TE * sizeof(float2) + 256;

struct InEdgeArgs {
    const float* e;            // [E][128] edge features, COO order
    float e_scale;             // 2^l
    const int32_t* perm;       // [E] COO edge id of every aggregation-order position (NULL: identity)
    const float* pq;           // [N][256]  P | Q
    const int32_t* rowptr;     // plan
    const int32_t* dstv;
    const int32_t* srcv;
    int64_t n_edges;
    const void* wimg;          // [5][hi | lo] images of We, W1..W4
    const float* bias;         // [5][128] (row 0 unused: b0 is folded into P)
    const float* gamma;        // LayerNorm affine
    const float* beta;
    float* agg;                // [N][128], pre-zeroed
    float* part_head;
    float* part_tail;
    int* range_flag;
};

// Last-layer phase B for one thread (= channel n of TMEM lane quadrant n / 32) and the 64 positions [hf*64, hf*64+64) of a
// tile: m = LayerNorm(y) from the accumulator (re-read) and the per-edge statistics, then the segmented mean over the
// destination-sorted positions — one pass per stored sum, warp-uniform control flow, no atomics (segmeta.cuh).
__device__ __forceinline__ void ie_norm_reduce(uint32_t tacc, const IeMeta* M, const float2* st, int hf, int n, float bias, float gamma,
                                               float beta) {
    float sum = 0.f;
#pragma unroll 1
    for (int cb = 0; cb < 64; cb += 8) {
        const int c0 = hf * 64 + cb;
        float v[8];
        umma::tmem_ld8(tacc + (uint32_t)cb, v);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float4 mr = *reinterpret_cast<const float4*>(st + c0 + i);     // mean, rstd of two edges
            v[i] = fmaf(((v[i] + bias) - mr.x) * mr.y, gamma, beta);
            v[i + 1] = fmaf(((v[i + 1] + bias) - mr.z) * mr.w, gamma, beta);
        }
        uint32_t fm = (M->flushmask[c0 >> 5] >> (c0 & 31)) & 0xffu, todo = 0xffu;
        if (fm == 0) {                   // the common case: no stored sum ends inside the chunk
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

#ifdef MGB_TIMELINE
__device__ long long* g_ie_timeline = nullptr;      // [role 0..3][pair 0..3][layer 0..4][tile 0..1][event 0..3]
#define IETL(role, it_, l_, t_, ev) do { if (blockIdx.x == 0 && (it_) < 4 && (threadIdx.x & 31) == 0 && g_ie_timeline) g_ie_timeline[((((role) * 4 + (it_)) * 5 + (l_)) * 2 + (t_)) * 4 + (ev)] = clock64(); } while (0)
int set_ie_timeline_buffer(long long* p) {
    return cudaMemcpyToSymbol(g_ie_timeline, &p, sizeof(p)) == cudaSuccess ? MGB_OK : MGB_ERR_CUDA;
}
#else
#define IETL(role, it_, l_, t_, ev) do { } while (0)
#endif

template <int NSPLIT>
__global__ void __launch_bounds__(IE_THREADS, 1) in_edge_fwd_tc_kernel(const InEdgeArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* w_img = base;                                        // [hi|lo] of the current layer
    unsigned char* x_img = base + (size_t)2 * TILE_BYTES;               // [tile 0|1][hi|lo]; fp32 staging in the last layer
    IeMeta* metas = reinterpret_cast<IeMeta*>(x_img + (size_t)4 * TILE_BYTES);
    float2* stats = reinterpret_cast<float2*>(metas + IE_MSLOTS);       // [pair parity][tile][128] mean, rstd of every edge
    uint64_t* bars = reinterpret_cast<uint64_t*>(stats + 4 * IE_TE);
    uint64_t* x_full = bars;          // [2] producers -> MMA (layer 0 operand written)
    // (the tile slots go back to the producers through hardware barriers 3 and 4: a waiting producer issues nothing)
    uint64_t* t_full = bars + 4;      // [2] MMA -> epilogue
    uint64_t* x_ready = bars + 6;     // [2] hidden-layer epilogue -> MMA
    uint64_t* w_bar = bars + 8;
    uint64_t* w_free = bars + 9;
    uint64_t* m_full = bars + 10;     // [IE_MSLOTS] meta warp -> epilogue
    uint64_t* m_empty = bars + 10 + IE_MSLOTS;   // [IE_MSLOTS] epilogue -> meta warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10 + 2 * IE_MSLOTS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_edges, IE_TE);
    const int64_t n_pairs = (n_tiles + 1) / 2;
    const int np = (int)((n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tile pairs of this CTA (>= 1)

    if (tid == 0) {
        for (int t = 0; t < 2; ++t) {
            umma::mbar_init(&x_full[t], IE_PROD_WARPS * 32);
            umma::mbar_init(&t_full[t], 1);
            umma::mbar_init(&x_ready[t], IE_EPI_WARPS * 32);
        }
        for (int s = 0; s < IE_MSLOTS; ++s) {
            umma::mbar_init(&m_full[s], 32);
            umma::mbar_init(&m_empty[s], IE_EPI_WARPS * 32);
        }
        umma::mbar_init(w_bar, 1);
        umma::mbar_init(w_free, 1);
        umma::fence_barrier_init();
    }
    if (warp == IE_MMA_WARP) umma::tmem_alloc(tmem_slot, 512);      // 4 accumulators: (pair parity, tile of the pair)
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < IE_EPI_WARPS) {
        // =========================== epilogue: thread = output channel n; warps 0-3 positions 0-63, warps 4-7 64-127 ====
        const int n = tid & 127, hf = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float gamma = a.gamma[n], beta = a.beta[n];
        uint32_t tf[2] = {0, 0};        // completed phases of t_full[t]
#pragma unroll 1
        for (int it = 0; it < np; ++it) {
            const int64_t pair = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
#pragma unroll 1
            for (int l = 0; l < IE_L - 1; ++l) {
                // ---- layers 0-3: bias, ReLU, next operand in place (MN-major image [n][e]); layer 0's accumulator was
                // initialised with P[dst] + Q[src] by the producers (b0 is folded into P)
                const float bias = l ? a.bias[l * 128 + n] : 0.f;
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    if (pair * 2 + t >= n_tiles) continue;
                    unsigned char* xrow = x_img + (size_t)t * 2 * TILE_BYTES + n * 128;
                    const uint32_t tacc = tmem + (uint32_t)(((it & 1) * 2 + t) * 128) + lane_base + (uint32_t)(hf * 64);
                    umma::mbar_wait(&t_full[t], tf[t] & 1);
                    ++tf[t];
                    umma::tc_fence_after();
                    if (warp == 0 || warp == 4) IETL(1 + hf, it, l, t, 0);
                    float vmax = 0.f;
#pragma unroll 1
                    for (int cb = 0; cb < 64; cb += 32) {
                        const int c0 = hf * 64 + cb;
                        float v[32];
                        umma::tmem_ld32(tacc + (uint32_t)cb, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bias, 0.f);
                        if (NSPLIT == 2) {
#pragma unroll
                            for (int i = 0; i < 32; i += 2) vmax = fmaxf(vmax, fmaxf(v[i], v[i + 1]));
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int cg = c0 + g * 8;
                            const uint32_t off = (uint32_t)(cg >> 6) * (128u * 128u) + (uint32_t)((((cg & 63) >> 3) ^ (n & 7)) << 4);
                            uint4 hi, lo;
                            if (NSPLIT == 2) {
                                split2_f16(v[g * 8 + 0], v[g * 8 + 1], hi.x, lo.x);
                                split2_f16(v[g * 8 + 2], v[g * 8 + 3], hi.y, lo.y);
                                split2_f16(v[g * 8 + 4], v[g * 8 + 5], hi.z, lo.z);
                                split2_f16(v[g * 8 + 6], v[g * 8 + 7], hi.w, lo.w);
                                *reinterpret_cast<uint4*>(xrow + off) = hi;
                                *reinterpret_cast<uint4*>(xrow + TILE_BYTES + off) = lo;
                            } else {
                                hi.x = umma::pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
                                hi.y = umma::pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
                                hi.z = umma::pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
                                hi.w = umma::pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
                                *reinterpret_cast<uint4*>(xrow + off) = hi;
                            }
                        }
                    }
                    if (NSPLIT == 2 && vmax >= 32768.f && a.range_flag) *a.range_flag = 1;
                    umma::fence_async_smem();
                    umma::tc_fence_before();
                    umma::mbar_arrive(&x_ready[t]);
                    if (warp == 0 || warp == 4) IETL(1 + hf, it, l, t, 1);
                }
            }
            // ---- last layer: y = D + b4 -> LayerNorm over channels -> segmented mean over positions.
            // Phase A (both tiles): the LayerNorm statistics go through the operand tile, which is handed back to the producers
            // right after; phase B (both tiles): normalise from TMEM + reduce — overlaps the next pair's first layers.
            const float bias = a.bias[(IE_L - 1) * 128 + n];
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                if (pair * 2 + t >= n_tiles) continue;
                unsigned char* xt = x_img + (size_t)t * 2 * TILE_BYTES;
                const uint32_t tacc = tmem + (uint32_t)(((it & 1) * 2 + t) * 128) + lane_base + (uint32_t)(hf * 64);
                umma::mbar_wait(&t_full[t], tf[t] & 1);
                ++tf[t];
                umma::tc_fence_after();
                if (warp == 0 || warp == 4) IETL(1 + hf, it, IE_L - 1, t, 0);
                // pass 1: y[n][e] -> staging [e][n] fp32 (16-byte chunks of a row permuted by the row index: every
                // access pattern below is bank-conflict free).  The operand tile is free: all MMAs that read it are done.
#pragma unroll 1
                for (int cb = 0; cb < 64; cb += 16) {
                    const int c0 = hf * 64 + cb;
                    float v[16];
                    umma::tmem_ld16(tacc + (uint32_t)cb, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int e = c0 + i;
                        *reinterpret_cast<float*>(xt + e * 512 + ((((n >> 2) ^ (e & 31))) << 4) + (n & 3) * 4) = v[i] + bias;
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + hf) : "memory");
                // pass 2: two threads per edge (64 channels each), exact two-pass mean / variance
                {
                    const int j = tid & 127;
                    const int e = hf * 64 + (j >> 1), part = j & 1;
                    const unsigned char* row = xt + e * 512;
                    float4 y[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int chunk = part * 16 + (i ^ (part << 2));
                        y[i] = *reinterpret_cast<const float4*>(row + ((chunk ^ (e & 31)) << 4));
                    }
                    float s = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) s += (y[i].x + y[i].y) + (y[i].z + y[i].w);
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    const float mean = s * (1.0f / 128.0f);
                    float q = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float dx = y[i].x - mean, dy = y[i].y - mean, dz = y[i].z - mean, dw = y[i].w - mean;
                        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                    }
                    q += __shfl_xor_sync(0xffffffffu, q, 1);
                    if (part == 0) stats[((it & 1) * 2 + t) * IE_TE + e] = make_float2(mean, 1.0f / sqrtf(q * (1.0f / 128.0f) + 1e-5f));
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + hf) : "memory");
                // the tile slot goes back to the producers (hardware barrier: the waiting side costs no issue slots), which refill
                // it while phase B runs
                asm volatile("bar.arrive %0, 512;" ::"r"(3 + t) : "memory");
                if (warp == 0 || warp == 4) IETL(1 + hf, it, IE_L - 1, t, 2);
            }
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                if (pair * 2 + t >= n_tiles) continue;
                const int slot = (it & 1) * 2 + t;
                umma::mbar_wait(&m_full[slot], (it >> 1) & 1);
                ie_norm_reduce(tmem + (uint32_t)(slot * 128) + lane_base + (uint32_t)(hf * 64), metas + slot, stats + slot * IE_TE, hf, n, bias,
                               gamma, beta);
                umma::tc_fence_before();
                umma::mbar_arrive(&m_empty[slot]);
                if (warp == 0 || warp == 4) IETL(1 + hf, it, IE_L - 1, t, 1);
            }
        }
    } else if (warp == IE_MMA_WARP) {
        // =========================== MMA issue + weight loads =======================================
        const uint32_t id_k = NSPLIT == 2 ? umma::idesc_f16(128, 128, 0, 0) : umma::idesc_bf16(128, 128, 0, 0);   // layer 0: B K-major
        const uint32_t id_m = NSPLIT == 2 ? umma::idesc_f16(128, 128, 0, 1) : umma::idesc_bf16(128, 128, 0, 1);   // layers >= 1: B MN-major
        const uint64_t w_d = umma::desc_sw128(umma::smem_u32(w_img), 16, 1024);
        const uint64_t xk_d = umma::desc_sw128(umma::smem_u32(x_img), 16, 1024);
        const uint64_t xm_d = umma::desc_sw128(umma::smem_u32(x_img), 128 * 128, 1024);
        constexpr uint32_t TB = TILE_BYTES >> 4;
        uint32_t wl = 0;                 // weight loads issued so far
        uint32_t xf[2] = {0, 0}, xr[2] = {0, 0};
#pragma unroll 1
        for (int it = 0; it < np; ++it) {
            const int64_t pair = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
#pragma unroll 1
            for (int l = 0; l < IE_L; ++l) {
                if (wl > 0) umma::mbar_wait(w_free, (wl - 1) & 1);          // the MMAs of the previous layer are done with w_img
                if (lane == 0) load_w2_image(w_img, (const unsigned char*)a.wimg + (size_t)l * 2 * TILE_BYTES, NSPLIT * TILE_BYTES, w_bar);
                IETL(0, it, l, 0, 3);
                umma::mbar_wait(w_bar, wl & 1);
                ++wl;
                IETL(0, it, l, 0, 0);
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    if (pair * 2 + t >= n_tiles) continue;
                    if (l == 0) { umma::mbar_wait(&x_full[t], xf[t] & 1); ++xf[t]; }
                    else { umma::mbar_wait(&x_ready[t], xr[t] & 1); ++xr[t]; }
                    umma::tc_fence_after();
                    IETL(0, it, l, t, 1);
                    if (umma::elect_one()) {
                        const uint32_t d = tmem + (uint32_t)(((it & 1) * 2 + t) * 128);
                        const uint64_t xd = (l == 0 ? xk_d : xm_d) + (uint64_t)((uint32_t)t * 2 * TB);
#pragma unroll
                        for (int term = 0; term < (NSPLIT == 2 ? 3 : 1); ++term) {      // small terms first: lo*hi, hi*lo, hi*hi
                            const uint64_t wa = w_d + ((NSPLIT == 2 && term == 0) ? TB : 0), xb = xd + ((NSPLIT == 2 && term == 1) ? TB : 0);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t koff_k = (uint32_t)((k >> 2) * (128 * 128 >> 4) + (k & 3) * 2), koff_m = (uint32_t)(k * 128);
                                umma::mma_bf16(d, wa + (uint64_t)koff_k, xb + (uint64_t)(l == 0 ? koff_k : koff_m), l == 0 ? id_k : id_m,
                                               (l == 0 || (term | k)) ? 1u : 0u);
                            }
                        }
                        umma::mma_commit(&t_full[t]);
                    }
                    __syncwarp();
                    IETL(0, it, l, t, 2);
                }
                if (umma::elect_one()) umma::mma_commit(w_free);
                __syncwarp();
            }
        }
    } else if (warp == IE_META_WARP) {
        // =========================== segment metadata, one pair ahead ===============================
#pragma unroll 1
        for (int it = 0; it < np; ++it) {
            const int64_t pair = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                if (pair * 2 + t >= n_tiles) continue;
                const int slot = (it & 1) * 2 + t;
                umma::mbar_wait_relaxed<2000>(&m_empty[slot], ((it >> 1) & 1) ^ 1);
                build_tile_meta(metas + slot, a.rowptr, a.dstv, a.srcv, a.n_edges, pair * 2 + t, lane, a.agg, SEG_H, true, a.part_head,
                                a.part_tail, IE_FLUSH);
                umma::mbar_arrive(&m_full[slot]);
            }
        }
    } else {
        // =========================== producers: layer-0 operand, accumulator initialisation, last-layer phase B ==========
        // (a) 16 e_features rows per warp -> fp16 hi | lo K-major image (gathered and converted BEFORE the tile slot is
        //     waited for);  (b) D^T[n][e] := P[dst_e][n] + Q[src_e][n] written straight into the accumulator with tcgen05.st
        //     (thread = channel n of the warp's TMEM lane quadrant, 64 positions per warp), so that layer 0's MMAs add
        //     We e on top and its epilogue is the plain bias-free ReLU — the gathers never sit on the epilogue's path;
        const int pw = warp - IE_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        const float* src = a.e + lane * 4;
        const float sc = a.e_scale;
        const int quad = warp & 3, chalf = pw >> 2;            // TMEM lane quadrant of this warp, its half of the positions
        const int n = quad * 32 + lane;
        const float* pqn = a.pq + n;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
#pragma unroll 1
        for (int it = 0; it <= np; ++it) {                      // iteration np only takes the last pair's slot hand-overs
            const int64_t pair = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            // (b) first, for both tiles of the pair: their accumulators were drained two pairs ago (the hand-over of the previous
            // pair's slot came after that), so they are initialised while the previous pair is still in flight
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                const int64_t tile = pair * 2 + t;
                if (it == np || tile >= n_tiles) continue;
                // the accumulator was last read by phase B of pair it - 2 (normally long done: the wait falls through)
                if (it >= 2) {
                    umma::mbar_wait_relaxed<200>(&m_empty[(it & 1) * 2 + t], ((it - 2) >> 1) & 1);
                    umma::tc_fence_after();
                }
                const int64_t c0 = tile * IE_TE + chalf * 64;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const int64_t p = c0 + h * 32 + lane;
                    // positions past the end of the edge list read node 0 (finite values in columns that belong to no segment);
                    // no select behind the loads: all 32 stay in flight
                    const uint32_t dl = p < a.n_edges ? (uint32_t)a.dstv[p] : 0u;
                    const uint32_t sl = p < a.n_edges ? (uint32_t)a.srcv[p] : 0u;
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = pqn[(size_t)__shfl_sync(0xffffffffu, sl, i) * 256 + 128];
#pragma unroll
                    for (int g = 0; g < 32; g += 8) {                    // P rows repeat along a segment: mostly L1 hits
                        float pv[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) pv[i] = pqn[(size_t)__shfl_sync(0xffffffffu, dl, g + i) * 256];
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[g + i] += pv[i];
                    }
                    umma::tmem_st32(tmem + (uint32_t)(((it & 1) * 2 + t) * 128) + lane_base + (uint32_t)(chalf * 64 + h * 32), acc);
                }
            }
            // the COO ids of this warp's 16 rows of both tiles; tile 1's rows are pulled into L2 now, read after tile 0 is stored
            int64_t mine[2] = {-1, -1};
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int64_t p0 = (pair * 2 + t) * IE_TE + pw * 16;
                if (it < np && lane < 16 && p0 + lane < a.n_edges) mine[t] = a.perm ? (int64_t)a.perm[p0 + lane] : p0 + lane;
            }
            if (mine[1] >= 0) {
                const float* row = a.e + mine[1] * 128;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 32));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 64));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 96));
            }
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                const bool cur = it < np && pair * 2 + t < n_tiles;                       // tile (it, t) exists
                const bool prev = it > 0 && (pair - gridDim.x) * 2 + t < n_tiles;          // tile (it - 1, t) exists
                uint4 hl[16];
                if (cur) {
                    // (a) e_features rows -> registers -> images once the tile slot is free
                    const int64_t mt = t ? mine[1] : mine[0];
                    float4 x[16];
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const int64_t c = __shfl_sync(0xffffffffu, mt, r);
                        x[r] = *reinterpret_cast<const float4*>(src + (c < 0 ? 0 : c) * 128);
                        if (c < 0) x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        float4 h = x[r];
                        h.x *= sc; h.y *= sc; h.z *= sc; h.w *= sc;
                        if (NSPLIT == 2) {
                            if (fmaxf(fmaxf(fabsf(h.x), fabsf(h.y)), fmaxf(fabsf(h.z), fabsf(h.w))) >= 32768.f && a.range_flag) *a.range_flag = 1;
                            split2_f16(h.x, h.y, hl[r].x, hl[r].z);
                            split2_f16(h.z, h.w, hl[r].y, hl[r].w);
                        } else {
                            hl[r].x = umma::pack_bf16(h.x, h.y);
                            hl[r].y = umma::pack_bf16(h.z, h.w);
                        }
                    }
                }
                if (pw == 0) IETL(3, it, 0, t, 0);
                if (prev) {
                    asm volatile("bar.sync %0, 512;" ::"r"(3 + t) : "memory");      // statistics of (it-1, t) written, slot t free
                    umma::tc_fence_after();
                }
                if (pw == 0) IETL(3, it, 0, t, 1);
                if (cur) {
                    unsigned char* img = x_img + (size_t)t * 2 * TILE_BYTES;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                        *reinterpret_cast<uint2*>(img + off) = make_uint2(hl[r].x, hl[r].y);
                        if (NSPLIT == 2) *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = make_uint2(hl[r].z, hl[r].w);
                    }
                    umma::fence_async_smem();
                    umma::tc_fence_before();
                    umma::mbar_arrive(&x_full[t]);
                }
                if (pw == 0) IETL(3, it, 0, t, 2);
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == IE_MMA_WARP) umma::tmem_dealloc(tmem, 512);
}

size_t in_edge_fwd_workspace(int64_t n_edges) {
    const int64_t sub = ceil_div<int64_t>(n_edges > 0 ? n_edges : 1, IE_FLUSH);
    return 2 * align_up((size_t)sub * SEG_H * sizeof(float)) + 512;
}

// precision: 2 = plain bf16 (1e-2 contract), anything else = fp16 hi/lo split (1e-5 contract)
int launch_in_edge_fwd(int precision, const float* e, float e_scale, const int32_t* perm, const float* pq, const int32_t* rowptr,
                       const int32_t* dstv, const int32_t* srcv, int64_t n_nodes, int64_t n_edges, const float* packed, float* agg,
                       int* range_flag, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(n_edges >= 0 && n_edges < ((int64_t)1 << 31) && n_nodes >= 0 && n_nodes < ((int64_t)1 << 23) * 256,
                "in_edge_fwd: sizes out of range");
    MGB_REQUIRE(((uintptr_t)e % 16) == 0 && ((uintptr_t)pq % 16) == 0, "in_edge_fwd: e / pq must be 16-byte aligned");
    if (n_nodes > 0) MGB_CUDA(cudaMemsetAsync(agg, 0, (size_t)n_nodes * SEG_H * sizeof(float), s));
    if (n_edges <= 0) return MGB_OK;
    const int64_t sub = ceil_div<int64_t>(n_edges, IE_FLUSH);
    Workspace ws(ws_ptr, ws_bytes);
    float* part_head = ws.take<float>((size_t)sub * SEG_H);
    float* part_tail = ws.take<float>((size_t)sub * SEG_H);
    MGB_WS_CHECK(ws);
    InEdgeArgs a{};
    a.e = e; a.e_scale = e_scale; a.perm = perm; a.pq = pq; a.rowptr = rowptr; a.dstv = dstv; a.srcv = srcv; a.n_edges = n_edges;
    a.wimg = packed;
    a.bias = packed + (size_t)IE_L * 128 * 128;
    a.gamma = a.bias + IE_L * 128;
    a.beta = a.gamma + 128;
    a.agg = agg; a.part_head = part_head; a.part_tail = part_tail; a.range_flag = range_flag;
    const int64_t pairs = (ceil_div<int64_t>(n_edges, IE_TE) + 1) / 2;
    const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
    {
        ProfScope prof(PROF_IN_EDGE_FWD, s);
        if (precision == 2) {
            MGB_CUDA(cudaFuncSetAttribute(in_edge_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IN_EDGE_SMEM));
            in_edge_fwd_tc_kernel<1><<<grid, IE_THREADS, IN_EDGE_SMEM, s>>>(a);
        } else {
            MGB_CUDA(cudaFuncSetAttribute(in_edge_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IN_EDGE_SMEM));
            in_edge_fwd_tc_kernel<2><<<grid, IE_THREADS, IN_EDGE_SMEM, s>>>(a);
        }
    }
    MGB_LAUNCH_CHECK();
    return launch_segment_fixup(rowptr, dstv, n_edges, IE_FLUSH, part_head, part_tail, agg, SEG_H, 1, s);
}

}  // namespace mgb
