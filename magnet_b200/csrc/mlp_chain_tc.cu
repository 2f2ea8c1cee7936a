// magnet_b200 — a whole MLP (models/backbones/mlp.py:9-28: Linear(128,128) + ReLU ... Linear(128,out)) in ONE kernel.
//
// The per-layer tensor-core Linear (linear_tc.cu) streams [rows,128] fp32 activations through HBM once per layer; the
// 4+1-layer MLPs of MAgNet (projector over every (query, t) row, the edge / node functions of the 10 InteractionNetworks)
// are then HBM-bound on traffic that never needs to leave the SM.  Here a CTA keeps TWO 128-row tiles of activations in
// shared memory (fp16 hi | lo images, 64 KB each) and walks the layers.  The WEIGHTS are the A operand and live in TENSOR
// MEMORY (TS form of tcgen05.mma: lane = output channel, packed fp16 pairs along K; two layer slots of 128 columns, filled
// one layer ahead by four loader warps with coalesced 16-byte loads + tcgen05.st): a 128x128 SS-MMA reads 8 KB of operands
// per 64 cycles = the whole 128 B/clk of shared memory; with A in TMEM it reads 4 KB, the weight re-load bubble between the
// layers disappears, and 64 KB of shared memory are free:
//   layer 0     : producers load x rows (fp32, optional ReLU on load) -> K-major images          (as linear_tc.cu)
//   every layer : D^T[n][row] = sum_k W_l[n][k] X_l[row][k]   (three fp16-split MMA terms, small terms first, TMEM)
//   hidden layer: epilogue (thread = channel n) adds bias, ReLU, splits into fp16 hi | lo and writes X_{l+1} IN PLACE
//                 as an MN-major image [n][row] — exactly the operand form the next layer's MMA reads (B MN-major)
//   last layer  : epilogue adds bias and stores y[row][n] (n < n_out), coalesced per row
// Two tiles ping-pong between the MMA warp and the epilogue warps while the producers fill a THIRD operand slot (possible
// because the weights left shared memory), so filling a tile never sits on the pipeline's critical path and the producers
// write their rows straight to shared memory.  Forward only (inference / rollout);
// same arithmetic as mgb_linear_tc_fwd with precision 3 (fp16 hi/lo split, 22 significant bits).
//
// MODE 1 — the fused INR decoder (MAgNetGNN.continuous_decoder + projector, models/magnet_gnn.py:224-283,339): the rows of
// layer 0 are not loaded but COMPUTED by the producers, row = (query q, time step i):
//   z[q,i,:] = blend of the proj_head outputs of the two nearest low-res nodes (latent part factorised per node: A; relative
//   coordinates, input value and time through the small proj_head columns), exactly the arithmetic of inr_decode_fwd_kernel;
// two SEARCH warps run a pair ahead (each owns every other pair): per query of the tiles they find the two nearest low-res nodes in the grid hash
// (knn_query, grid.cuh — or read a precomputed neighbour table) and leave a 48-byte record (nodes, relative coordinates,
// blend weights) in shared memory.  z [Q,T,128] (5 KB per query) never exists: HBM sees the query coordinates in
// and hr_points [Q,T] out, plus L2-resident gathers of A.
#include "internal.cuh"
#include "tc_common.cuh"
#include "grid.cuh"

namespace mgb {

// warps (whole warpgroups, so that registers can follow the work): 0-7 epilogue, 8 MMA issue, 9-12 weight loaders (TMEM lane
// quadrants 1,2,3,0), 13-14 search (MODE 1), 15 spare, 16-23 producers
constexpr int MC_EPI_WARPS = 8, MC_PROD_WARPS = 8;
constexpr int MC_MMA_WARP = 8, MC_W_WARP0 = 9, MC_SEARCH_WARP0 = 13;
template <int MODE> constexpr int mc_prod_warp0() { return 16; }
template <int MODE> constexpr int mc_threads() { return (mc_prod_warp0<MODE>() + MC_PROD_WARPS) * 32; }      // 768
constexpr int MC_SLOTS = 3;             // operand tile slots: two tiles in the MMA / epilogue pipeline, one being filled
constexpr uint32_t MC_TMEM_W = 256;      // first TMEM column of the weight slots (accumulators: columns 0..255)
constexpr int MC_MAXQ = 132;          // queries touched by one 128-row tile: at most 128 / T + 2
struct alignas(16) InrQRec { int sel0, sel1; float r0x, r0y, r1x, r1y, m0, m1, den; int b; int pad0, pad1; };
template <int MODE> constexpr size_t mlp_chain_smem() {
    return 1024 + (size_t)MC_SLOTS * 2 * TILE_BYTES + (MODE ? 4 * MC_MAXQ * sizeof(InrQRec) : 0) + 256;
}

// the two nearest low-res nodes of query q and the blend of MAgNetGNN.continuous_decoder (models/magnet_gnn.py:259-279):
// m0 / m1 multiply lat_0 / lat_1, den = w_1 + w_0 (same operations as inr_weights, interaction.cu)
template <int D>
__device__ __forceinline__ void inr_build_record(const InrFuseArgs& f, const GridParams& gp, int64_t q, InrQRec* rec) {
    const int b = (int)(q / f.nq_per_sample);
    const float hx = f.hr_coords[q * D], hy = D > 1 ? f.hr_coords[q * D + 1] : 0.f;
    int sel0, sel1;
    if (f.idx) {
        sel0 = (int)f.idx[q * f.k];
        sel1 = (int)f.idx[q * f.k + 1];
    } else {
        BestList<2> best;
        knn_query<2, D>(f.pts, f.cell_start, gp, hx, hy, b, (int64_t)f.L, 2, best);
        sel0 = best.i[0];
        sel1 = best.i[1];
    }
    InrQRec r;
    r.sel0 = sel0; r.sel1 = sel1; r.b = b; r.pad0 = r.pad1 = 0;
    r.r0x = f.lr_coords[(int64_t)sel0 * D] - hx;
    r.r1x = f.lr_coords[(int64_t)sel1 * D] - hx;
    r.r0y = D > 1 ? f.lr_coords[(int64_t)sel0 * D + 1] - hy : 0.f;
    r.r1y = D > 1 ? f.lr_coords[(int64_t)sel1 * D + 1] - hy : 0.f;
    float s0 = r.r0x * r.r0x, s1 = r.r1x * r.r1x;
    if (D > 1) { s0 += r.r0y * r.r0y; s1 += r.r1y * r.r1y; }
    const float n0 = sqrtf(s0), n1 = sqrtf(s1);      // the reference squares torch.norm(...)
    float w0 = n0 * n0, w1 = n1 * n1;
    if (f.mode == 1) { w0 = 1.0f / w0; w1 = 1.0f / w1; }
    if (f.mode == 2) { const float b0 = 1.0f - (float)f.L * w0, b1 = 1.0f - (float)f.L * w1; w0 = b0 * b0 * b0; w1 = b1 * b1 * b1; }
    r.den = w1 + w0;
    // the blend weights, already divided: z = lat_0 (m_0 / den) + lat_1 (m_1 / den) — one rounding more than the reference's
    // (lat_0 m_0 + lat_1 m_1) / den (1e-7), sixteen instructions less per channel
    if (f.mode == 0) { r.m0 = __fdiv_rn(w1, r.den); r.m1 = __fdiv_rn(w0, r.den); }       // area: neighbour 0 is weighted by the OTHER one's area
    else { r.m0 = __fdiv_rn(w0, r.den); r.m1 = __fdiv_rn(w1, r.den); }
    *rec = r;
}

template <int MODE>
__global__ void __launch_bounds__(mc_threads<MODE>(), 1) mlp_chain_tc_kernel(const MlpChainArgs a) {
    constexpr int MC_PROD_WARP0 = mc_prod_warp0<MODE>();
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* x_img = base;                                        // [slot][hi|lo]; tile j of this CTA lives in slot j % 3
    uint64_t* bars = reinterpret_cast<uint64_t*>(x_img + (size_t)MC_SLOTS * 2 * TILE_BYTES);
    uint64_t* x_full = bars;          // [3] producers -> MMA (layer-0 operand of the slot written)
    uint64_t* acc_free = bars + 3;    // [2] last-layer epilogue -> MMA (accumulator t drained)
    uint64_t* t_full = bars + 5;      // [2] MMA -> epilogue (accumulator of the layer ready)
    uint64_t* x_ready = bars + 7;     // [2] hidden-layer epilogue -> MMA (next operand written, accumulator drained)
    uint64_t* w_full = bars + 9;      // [2] loader warps -> MMA (weights of a layer in their TMEM slot)
    uint64_t* w_free = bars + 11;     // [2] MMA -> loader warps (every MMA that reads the slot has completed)
    uint64_t* q_full = bars + 13;     // [4] MODE 1: search warp -> producers (records of the tile written)
    uint64_t* q_empty = bars + 17;    // [4] MODE 1: producers -> search warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
    InrQRec* qrec = reinterpret_cast<InrQRec*>(bars + 22);      // [4][MC_MAXQ]
    // (the operand slots go back to the producers through hardware barriers 1..3: a waiting producer issues nothing)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.rows, 128);
    const int64_t n_pairs = (n_tiles + 1) / 2;
    const int np = (int)((n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tile pairs of this CTA (>= 1)
    // tiles of this CTA in processing order: j = 2 it + t; only the very last pair of the grid can lack its second tile
    const int nj = 2 * np - ((((int64_t)blockIdx.x + (int64_t)(np - 1) * gridDim.x) * 2 + 1 >= n_tiles) ? 1 : 0);
    const int L = a.n_layers;

    if (tid == 0) {
        for (int q = 0; q < MC_SLOTS; ++q) umma::mbar_init(&x_full[q], MC_PROD_WARPS * 32);
        for (int t = 0; t < 2; ++t) {
            umma::mbar_init(&acc_free[t], MC_EPI_WARPS * 32);
            umma::mbar_init(&t_full[t], 1);
            umma::mbar_init(&x_ready[t], MC_EPI_WARPS * 32);
            umma::mbar_init(&w_full[t], 4 * 32);
            umma::mbar_init(&w_free[t], 1);
        }
        for (int q = 0; q < 4; ++q) {
            umma::mbar_init(&q_full[q], 32);
            umma::mbar_init(&q_empty[q], MC_PROD_WARPS * 32);
        }
        umma::fence_barrier_init();
    }
    if (warp == MC_MMA_WARP) umma::tmem_alloc(tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < MC_EPI_WARPS) {
        // =========================== epilogue: thread = output channel n; warps 0-3 rows 0-63, warps 4-7 rows 64-127 ====
        const int n = tid & 127, hf = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t tf[2] = {0, 0};        // completed phases of t_full[t]
#pragma unroll 1
        for (int it = 0; it < np; ++it) {
            const int64_t pair = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const float bias = a.bias[l * 128 + n];
                const bool last = l == L - 1;
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    const int64_t tile = pair * 2 + t;
                    if (tile >= n_tiles) continue;
                    const int j = 2 * it + t, slot = j % MC_SLOTS;
                    const int64_t r0 = tile * 128;
                    const int nr = (int)((a.rows - r0) < 128 ? (a.rows - r0) : 128);
                    unsigned char* xrow = x_img + (size_t)slot * 2 * TILE_BYTES + n * 128;
                    umma::mbar_wait(&t_full[t], tf[t] & 1);
                    ++tf[t];
                    umma::tc_fence_after();
                    const uint32_t tacc = tmem + (uint32_t)(t * 128) + lane_base + (uint32_t)(hf * 64);
                    if (!last) {
                        // hidden layer: bias, activation, next operand in place (MN-major image [n][row]); 32 rows per step
                        float vmax = 0.f;
#pragma unroll 1
                        for (int cb = 0; cb < 64; cb += 32) {
                            const int c0 = hf * 64 + cb;
                            float v[32];
                            umma::tmem_ld32(tacc + (uint32_t)cb, v);
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] += bias;
                            if (a.act == ACT_RELU) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                            } else if (a.act == ACT_SWISH) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = swish_tc<false>(v[i]);
                            }
#pragma unroll
                            for (int i = 0; i < 32; i += 2) vmax = fmaxf(vmax, fmaxf(fabsf(v[i]), fabsf(v[i + 1])));
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                // X_{l+1}[row = cg .. cg+7][k = n] as an MN-major image: row n of the image, 16 bytes
                                const int cg = c0 + g * 8;
                                const uint32_t off = (uint32_t)(cg >> 6) * (128u * 128u) + (uint32_t)((((cg & 63) >> 3) ^ (n & 7)) << 4);
                                uint4 hi, lo;
                                split2_f16(v[g * 8 + 0], v[g * 8 + 1], hi.x, lo.x);
                                split2_f16(v[g * 8 + 2], v[g * 8 + 3], hi.y, lo.y);
                                split2_f16(v[g * 8 + 4], v[g * 8 + 5], hi.z, lo.z);
                                split2_f16(v[g * 8 + 6], v[g * 8 + 7], hi.w, lo.w);
                                *reinterpret_cast<uint4*>(xrow + off) = hi;
                                *reinterpret_cast<uint4*>(xrow + TILE_BYTES + off) = lo;
                            }
                        }
                        if (vmax >= 32768.f && a.range_flag) *a.range_flag = 1;
                        umma::fence_async_smem();
                        umma::tc_fence_before();
                        umma::mbar_arrive(&x_ready[t]);
                    } else {
                        // every MMA that read this tile's slot has completed: it goes back to the producers (hardware barrier),
                        // if a later tile of this CTA will use it
                        if (j + MC_SLOTS < nj) asm volatile("bar.arrive %0, 512;" ::"r"(1 + slot) : "memory");
                        // last layer: warps whose 32 channels lie beyond n_out have nothing to store (n_out = 1: the projector)
                        if ((warp & 3) * 32 < a.n_out) {
#pragma unroll 1
                            for (int cb = 0; cb < 64; cb += 8) {
                                const int c0 = hf * 64 + cb;
                                float v[8];
                                umma::tmem_ld8(tacc + (uint32_t)cb, v);
                                if (n < a.n_out) {
                                    float* yo = a.y + (r0 + c0) * a.ldy + n;
                                    const int lim = nr - c0;
#pragma unroll
                                    for (int i = 0; i < 8; ++i)
                                        if (i < lim) yo[(int64_t)i * a.ldy] = v[i] + bias;
                                }
                            }
                        }
                        umma::tc_fence_before();
                        umma::mbar_arrive(&acc_free[t]);
                    }
                }
            }
        }
    } else if (warp == MC_MMA_WARP) {
        // =========================== MMA issue =======================================================
        umma::reg_dec<56>();
        const uint32_t id_k = umma::idesc_f16(128, 128, 0, 0);       // layer 0: B K-major (rows of x)
        const uint32_t id_m = umma::idesc_f16(128, 128, 0, 1);       // layers >= 1: B MN-major ([n][row] images)
        const uint64_t xk_d = umma::desc_sw128(umma::smem_u32(x_img), 16, 1024);
        const uint64_t xm_d = umma::desc_sw128(umma::smem_u32(x_img), 128 * 128, 1024);
        constexpr uint32_t TB = TILE_BYTES >> 4;
        uint32_t wl = 0;                 // layers issued so far (their weights alternate between the two TMEM slots)
        uint32_t xr[2] = {0, 0};
#pragma unroll 1
        for (int it = 0; it < np; ++it) {
            const int64_t pair = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const uint32_t ws = wl & 1;
                umma::mbar_wait(&w_full[ws], (wl >> 1) & 1);
                const uint32_t w_t = tmem + MC_TMEM_W + ws * 128u;          // hi pairs: +0..63, lo pairs: +64..127
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    if (pair * 2 + t >= n_tiles) continue;
                    const int j = 2 * it + t, slot = j % MC_SLOTS;
                    if (l == 0) {
                        if (it > 0) umma::mbar_wait(&acc_free[t], (it - 1) & 1);        // accumulator t drained by the tile two back
                        umma::mbar_wait(&x_full[slot], (j / MC_SLOTS) & 1);
                    } else {
                        umma::mbar_wait(&x_ready[t], xr[t] & 1);
                        ++xr[t];
                    }
                    umma::tc_fence_after();
                    if (umma::elect_one()) {
                        const uint32_t d = tmem + (uint32_t)(t * 128);
                        const uint64_t xd = (l == 0 ? xk_d : xm_d) + (uint64_t)((uint32_t)slot * 2 * TB);
#pragma unroll
                        for (int term = 0; term < 3; ++term) {      // small terms first: lo*hi, hi*lo, hi*hi
                            const uint32_t wa = w_t + (term == 0 ? 64u : 0u);
                            const uint64_t xb = xd + (term == 1 ? TB : 0);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t koff_k = (uint32_t)((k >> 2) * (128 * 128 >> 4) + (k & 3) * 2), koff_m = (uint32_t)(k * 128);
                                umma::mma_bf16_ts(d, wa + (uint32_t)(k * 8), xb + (uint64_t)(l == 0 ? koff_k : koff_m), l == 0 ? id_k : id_m,
                                                  (term | k) ? 1u : 0u);
                            }
                        }
                        umma::mma_commit(&t_full[t]);
                    }
                    __syncwarp();
                }
                if (umma::elect_one()) umma::mma_commit(&w_free[ws]);
                __syncwarp();
                ++wl;
            }
        }
    } else if (warp >= MC_W_WARP0 && warp < MC_W_WARP0 + 4) {
        // =========================== weight loaders: one TMEM lane quadrant each, one layer ahead =================
        umma::reg_dec<56>();
        const uint32_t lane_q = (uint32_t)((warp & 3) * 32) << 16;
        const int m = (warp & 3) * 32 + lane;                          // output channel = TMEM lane = row of W_l
        uint32_t wl = 0;
#pragma unroll 1
        for (int it = 0; it < np; ++it) {
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const uint32_t ws = wl & 1;
                if (wl >= 2) umma::mbar_wait_relaxed<100>(&w_free[ws], ((wl >> 1) - 1) & 1);      // MMAs of the layer before last are done
                umma::tc_fence_after();
                const uint4* src = reinterpret_cast<const uint4*>(a.wimg) + (size_t)l * 4096 + m;
#pragma unroll 1
                for (int pc = 0; pc < 4; ++pc) {                        // (hi | lo) x (K pairs 0-31 | 32-63)
                    uint4 v[8];
#pragma unroll
                    for (int w4 = 0; w4 < 8; ++w4) v[w4] = src[(pc * 8 + w4) * 128];
                    float f[32];
#pragma unroll
                    for (int w4 = 0; w4 < 8; ++w4) {
                        f[w4 * 4 + 0] = __uint_as_float(v[w4].x); f[w4 * 4 + 1] = __uint_as_float(v[w4].y);
                        f[w4 * 4 + 2] = __uint_as_float(v[w4].z); f[w4 * 4 + 3] = __uint_as_float(v[w4].w);
                    }
                    umma::tmem_st32(tmem + lane_q + MC_TMEM_W + ws * 128u + (uint32_t)(pc * 32), f);
                }
                umma::tc_fence_before();
                umma::mbar_arrive(&w_full[ws]);
                ++wl;
            }
        }
    } else if (warp < MC_PROD_WARP0 && !(MODE == 1 && warp >= MC_SEARCH_WARP0 && warp < MC_SEARCH_WARP0 + 2)) {
        umma::reg_dec<56>();          // spare warps of the loader / search warpgroup
    } else if (MODE == 1 && warp < MC_PROD_WARP0) {
        // =========================== search warps: warp 13 -> both tiles of the even pairs, warp 14 -> of the odd pairs =====
        // A query's ring search is a chain of dependent L2 loads (~20 us) whatever the lane count, so a warp that owned one tile
        // of EVERY pair was exactly as slow as the pair itself (search-bound: 6.5 vs 5.2 ms per 2^20 queries).  Each warp now
        // takes the queries of BOTH tiles of every other pair side by side in its lanes and has two pair periods to do so.
        umma::reg_dec<56>();
        const InrFuseArgs& f = a.inr;
        const int w = warp - MC_SEARCH_WARP0;
        const GridParams gp = f.idx ? GridParams{} : *f.gp;
#pragma unroll 1
        for (int it = w; it < np; it += 2) {
            int64_t q0[2];
            int nq[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int64_t tile = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 2 + t;
                nq[t] = 0;
                q0[t] = 0;
                if (tile >= n_tiles) continue;
                umma::mbar_wait_relaxed<500>(&q_empty[(it & 1) * 2 + t], ((it >> 1) & 1) ^ 1);
                const int64_t r0 = tile * 128;
                const int64_t r1 = (r0 + 127 < a.rows - 1) ? r0 + 127 : a.rows - 1;
                q0[t] = r0 / f.T;
                nq[t] = (int)(r1 / f.T - q0[t]) + 1;
            }
            for (int j = lane; j < nq[0] + nq[1]; j += 32) {
                const int t = j < nq[0] ? 0 : 1, jq = t ? j - nq[0] : j;
                InrQRec* rec = qrec + ((it & 1) * 2 + t) * MC_MAXQ + jq;
                if (f.d == 2) inr_build_record<2>(f, gp, q0[t] + jq, rec);
                else inr_build_record<1>(f, gp, q0[t] + jq, rec);
            }
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (nq[t] > 0) umma::mbar_arrive(&q_full[(it & 1) * 2 + t]);
        }
    } else if (MODE == 1) {
        // =========================== producers (fused decoder): 16 rows = (query, time step) pairs per warp =========
        // rows go straight to the operand slot (it is free long before it is needed); the gathers of 4 rows are in flight together
        umma::reg_inc<104>();
        const InrFuseArgs& f = a.inr;
        const int pw = warp - MC_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        const int c = lane * 4;
        float wu[4], wt[4], wr0[4], wr1[4];          // the small proj_head columns of this lane's four channels
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float* w = f.wsmall + (int64_t)(c + u) * f.ldw;
            wu[u] = w[0];
            wr0[u] = w[1];
            wr1[u] = f.d > 1 ? w[2] : 0.f;
            wt[u] = w[1 + f.d];
        }
        const uint32_t T = (uint32_t)f.T;
#pragma unroll 1
        for (int j = 0; j < nj; ++j) {
            const int it = j >> 1, t = j & 1, slot = j % MC_SLOTS, qs = (it & 1) * 2 + t;
            const int64_t tile = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 2 + t;
            const uint32_t row_t = (uint32_t)(tile * 128);            // rows < 2^31
            const uint32_t q0 = row_t / T;
            umma::mbar_wait_relaxed<200>(&q_full[qs], (it >> 1) & 1);
            if (j >= MC_SLOTS) asm volatile("bar.sync %0, 512;" ::"r"(1 + slot) : "memory");      // slot handed back by the epilogue
            unsigned char* img = x_img + (size_t)slot * 2 * TILE_BYTES;
            const InrQRec* recs = qrec + qs * MC_MAXQ;
#pragma unroll 1
            for (int h = 0; h < 4; ++h) {
                const uint32_t rw = row_t + (uint32_t)(pw * 16 + h * 4);
                uint32_t q = rw / T, i = rw - q * T;                  // (query, time step) of the first of the 4 rows
                float4 A0[4], A1[4];
                float x0[4], x1[4], ti[4], m0[4], m1[4], rx0[4], ry0[4], rx1[4], ry1[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const bool ok = rw + r < (uint32_t)a.rows;
                    const InrQRec rec = recs[ok ? q - q0 : 0];
                    A0[r] = *reinterpret_cast<const float4*>(f.A + (int64_t)rec.sel0 * 128 + c);
                    A1[r] = *reinterpret_cast<const float4*>(f.A + (int64_t)rec.sel1 * 128 + c);
                    const float* xr = f.xlr + ((int64_t)rec.b * T + (ok ? i : 0)) * f.L - (int64_t)rec.b * f.L;
                    x0[r] = xr[rec.sel0];
                    x1[r] = xr[rec.sel1];
                    ti[r] = f.t[(int64_t)rec.b * f.ldt + (ok ? i : 0)];
                    m0[r] = ok ? rec.m0 : 0.f; m1[r] = ok ? rec.m1 : 0.f;         // (already divided by w_1 + w_0)
                    rx0[r] = rec.r0x; ry0[r] = rec.r0y; rx1[r] = rec.r1x; ry1[r] = rec.r1y;
                    if (++i == T) { i = 0; ++q; }
                }
                float zmax = 0.f;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float a0[4] = {A0[r].x, A0[r].y, A0[r].z, A0[r].w}, a1[4] = {A1[r].x, A1[r].y, A1[r].z, A1[r].w};
                    float z[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float c0 = fmaf(wr0[u], rx0[r], a0[u]), c1 = fmaf(wr0[u], rx1[r], a1[u]);
                        if (f.d > 1) { c0 = fmaf(wr1[u], ry0[r], c0); c1 = fmaf(wr1[u], ry1[r], c1); }
                        const float lat0 = fmaf(wt[u], ti[r], fmaf(wu[u], x0[r], c0));
                        const float lat1 = fmaf(wt[u], ti[r], fmaf(wu[u], x1[r], c1));
                        z[u] = fmaf(lat0, m0[r], lat1 * m1[r]);       // (lat_0 m_0 + lat_1 m_1) / (w_1 + w_0), weights pre-divided
                        zmax = fmaxf(zmax, fabsf(z[u]));
                    }
                    uint32_t h0, h1, l0, l1;
                    split2_f16(z[0], z[1], h0, l0);
                    split2_f16(z[2], z[3], h1, l1);
                    const int rr = h * 4 + r;
                    const uint32_t off = lane_blk + (uint32_t)(pw * 16 + rr) * 128u + ((lane_chunk ^ (uint32_t)(rr & 7)) << 4);
                    *reinterpret_cast<uint2*>(img + off) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = make_uint2(l0, l1);
                }
                if (zmax >= 32768.f && a.range_flag) *a.range_flag = 1;
            }
            umma::mbar_arrive(&q_empty[qs]);
            umma::fence_async_smem();
            umma::mbar_arrive(&x_full[slot]);
        }
    } else {
        // =========================== producers: 16 rows per warp (layer-0 operand), straight into the slot ==========
        umma::reg_inc<104>();
        const int pw = warp - MC_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        const float* src = a.x + lane * 4;
#pragma unroll 1
        for (int j = 0; j < nj; ++j) {
            const int it = j >> 1, t = j & 1, slot = j % MC_SLOTS;
            const int64_t tile = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 2 + t;
            const int64_t r0 = tile * 128 + pw * 16;
            float4 x[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int64_t row = r0 + r < a.rows ? r0 + r : a.rows - 1;
                x[r] = *reinterpret_cast<const float4*>(src + row * a.ldx);
            }
            if (j >= MC_SLOTS) asm volatile("bar.sync %0, 512;" ::"r"(1 + slot) : "memory");      // slot handed back by the epilogue
            unsigned char* img = x_img + (size_t)slot * 2 * TILE_BYTES;
            float xmax = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                float4 h = x[r];
                if (a.in_act == ACT_RELU) { h.x = fmaxf(h.x, 0.f); h.y = fmaxf(h.y, 0.f); h.z = fmaxf(h.z, 0.f); h.w = fmaxf(h.w, 0.f); }
                if (r0 + r >= a.rows) h = make_float4(0.f, 0.f, 0.f, 0.f);
                xmax = fmaxf(xmax, fmaxf(fmaxf(fabsf(h.x), fabsf(h.y)), fmaxf(fabsf(h.z), fabsf(h.w))));
                uint32_t h0, h1, l0, l1;
                split2_f16(h.x, h.y, h0, l0);
                split2_f16(h.z, h.w, h1, l1);
                const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                *reinterpret_cast<uint2*>(img + off) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = make_uint2(l0, l1);
            }
            if (xmax >= 32768.f && a.range_flag) *a.range_flag = 1;
            umma::fence_async_smem();
            umma::mbar_arrive(&x_full[slot]);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == MC_MMA_WARP) umma::tmem_dealloc(tmem, 512);
}

int launch_mlp_chain_tc(const MlpChainArgs& a, cudaStream_t s) {
    MGB_REQUIRE(a.n_layers >= 1 && a.n_layers <= 8, "mlp_chain: 1..8 layers (got %d)", a.n_layers);
    MGB_REQUIRE(a.n_out >= 1 && a.n_out <= 128, "mlp_chain: 1 <= n_out <= 128 (got %d)", a.n_out);
    MGB_REQUIRE(a.rows >= 0 && a.rows < ((int64_t)1 << 31), "mlp_chain: row count out of range");
    MGB_REQUIRE(a.ldx % 4 == 0 && ((uintptr_t)a.x % 16) == 0, "mlp_chain: x rows must be 16-byte aligned");
    if (a.rows == 0) return MGB_OK;
    const int64_t pairs = (ceil_div<int64_t>(a.rows, 128) + 1) / 2;
    const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
    MGB_CUDA(cudaFuncSetAttribute(mlp_chain_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp_chain_smem<0>()));
    ProfScope prof(PROF_NODE_GEMM, s);
    mlp_chain_tc_kernel<0><<<grid, mc_threads<0>(), mlp_chain_smem<0>(), s>>>(a);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

int launch_inr_decode_fused(const MlpChainArgs& a, cudaStream_t s) {
    const InrFuseArgs& f = a.inr;
    MGB_REQUIRE(a.n_layers >= 1 && a.n_layers <= 8 && a.n_out >= 1 && a.n_out <= 128, "inr_decode_fused: bad projector shape");
    MGB_REQUIRE(f.d == 1 || f.d == 2, "inr_decode_fused: coordinate dimension must be 1 or 2");
    MGB_REQUIRE(f.mode >= 0 && f.mode <= 2, "inr_decode_fused: unknown interpolation mode");
    MGB_REQUIRE(f.T >= 1 && f.L >= 2 && f.nq_per_sample >= 1 && (f.idx == nullptr || f.k >= 2), "inr_decode_fused: bad sizes");
    MGB_REQUIRE(a.rows == f.n_query * f.T && a.rows < ((int64_t)1 << 31), "inr_decode_fused: rows must be n_query * T (< 2^31)");
    MGB_REQUIRE(((uintptr_t)f.A % 16) == 0, "inr_decode_fused: A must be 16-byte aligned");
    if (a.rows == 0) return MGB_OK;
    const int64_t pairs = (ceil_div<int64_t>(a.rows, 128) + 1) / 2;
    const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
    MGB_CUDA(cudaFuncSetAttribute(mlp_chain_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp_chain_smem<1>()));
    ProfScope prof(PROF_INR_DECODE, s);
    mlp_chain_tc_kernel<1><<<grid, mc_threads<1>(), mlp_chain_smem<1>(), s>>>(a);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
