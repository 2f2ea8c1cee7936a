// magnet_b200 — the node update of GNN_Layer in ONE kernel (forward):
//     y1_pre = [x, agg, var] W3^T + b3          (update_net_1, models/mpnn_2d.py:57-60, :85-87)
//     y2_pre = Swish(y1_pre) W4^T + b4          (update_net_2, :61-62, :88)
//     out    = x + Swish(y2_pre)                (residual, :89;  InstanceNorm follows in gnn_layer.cu)
// As two launches of linear_tc.cu the intermediate [N,128] activation went to HBM and came back, x was read a second time
// as the residual, and each launch paid its own ramp (first tile: ~4 us before the first store) — the launches ran at about
// half of the HBM bound.  Here a CTA keeps a 128-row tile on the SM for both Linears:
//   * the WEIGHTS (W3 columns [0,128) and [128,256), W4; bf16 hi | lo) live in TENSOR MEMORY for the whole kernel as the
//     A operand of TS-form tcgen05.mma (lane = output channel, packed bf16 pairs along K; 3 x 128 columns), loaded once
//     per CTA by the epilogue warps; the remaining 128 columns are the fp32 accumulator;
//   * shared memory holds three 64 KB operand slots: producers write the x tile and the agg tile (K-major hi | lo images),
//     the epilogue of the first Linear writes Swish(y1_pre) IN PLACE over the agg slot as an MN-major image [k][row],
//     which is the B-operand form of the second Linear (same trick as mlp_chain_tc.cu);
//   * the epilogue thread owns one output channel: bias, the var column(s) of W3 (FFMA), stores of y1_pre / y2_pre (saved
//     for the backward pass), Swish, residual, out.
// HBM traffic per node: x, agg in (1 KB), y1_pre, y2_pre, out (1.5 KB) and the residual re-read of x (L2).
#include "internal.cuh"
#include "tc_common.cuh"

namespace mgb {

// warps (whole warpgroups: registers follow the work): 0-7 epilogue, 8 MMA issue, 9-11 idle, 12-19 producers
constexpr int NU_EPI_WARPS = 8, NU_PROD_WARPS = 8, NU_MMA_WARP = 8, NU_PROD_WARP0 = 12;
constexpr int NU_THREADS = (NU_PROD_WARP0 + NU_PROD_WARPS) * 32;      // 640
constexpr int NU_SLOTS = 3;
constexpr uint32_t NU_TMEM_W = 128;      // first TMEM column of the weights (accumulator: columns 0..127)
constexpr int NU_MAXV = 4;               // var columns handled in the epilogue

#ifdef MGB_TIMELINE
__device__ long long* g_nu_timeline = nullptr;      // [role 0..2][tile 0..7][event 0..3] of CTA 0 (tools/dev_nu_timeline.py)
#define NUTL(role, it_, ev) do { if (blockIdx.x == 0 && (it_) < 8 && (threadIdx.x & 31) == 0 && g_nu_timeline) g_nu_timeline[((role) * 8 + (it_)) * 4 + (ev)] = clock64(); } while (0)
int set_nu_timeline_buffer(long long* p) {
    return cudaMemcpyToSymbol(g_nu_timeline, &p, sizeof(p)) == cudaSuccess ? MGB_OK : MGB_ERR_CUDA;
}
#else
#define NUTL(role, it_, ev) do { } while (0)
#endif

template <int NSPLIT> constexpr size_t node_update_smem() { return 1024 + (size_t)NU_SLOTS * NSPLIT * TILE_BYTES + 256; }

// W [128][ld] (columns c0 .. c0+127) -> bf16 hi | lo pairs in the order the loader warps read it: 32-bit word j of row m =
// elements (2j, 2j+1); chunk (j / 32) x w4 ((j % 32) / 4) x m x 4 words (coalesced 16-byte loads per thread m); 8192 words
// hi, then 8192 words lo (same layout as pack_weight_tmem_kernel, linear_tc.cu, with bf16 instead of fp16 halves)
__global__ void pack_weight_tmem_bf16_kernel(const float* __restrict__ W, int ld, int c0, int trans, uint32_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * 64) return;
    const int m = idx >> 6, j = idx & 63;
    // trans = 0: A[m][k] = W[m][c0 + k];  trans = 1: A[m][k] = W[k][c0 + m] (the block of W^T: data gradients)
    const float v0 = trans ? W[(int64_t)(2 * j) * ld + c0 + m] : W[(int64_t)m * ld + c0 + 2 * j];
    const float v1 = trans ? W[(int64_t)(2 * j + 1) * ld + c0 + m] : W[(int64_t)m * ld + c0 + 2 * j + 1];
    uint32_t hi, lo;
    split2_bf16(v0, v1, hi, lo);
    const int word = (((j >> 5) * 8 + ((j & 31) >> 2)) * 128 + m) * 4 + (j & 3);
    out[word] = hi;
    out[8192 + word] = lo;
}

int pack_weight_tmem_bf16(const float* W, int ld, int c0, void* out, cudaStream_t s, int trans) {
    pack_weight_tmem_bf16_kernel<<<32, 256, 0, s>>>(W, ld, c0, trans, (uint32_t*)out);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

template <int NSPLIT, bool FAST>
__global__ void __launch_bounds__(NU_THREADS, 1) node_update_tc_kernel(const NodeUpdateArgs a) {
    constexpr uint32_t SLOT_BYTES = NSPLIT * TILE_BYTES;
    constexpr int NTERM = NSPLIT == 1 ? 1 : 3;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* x_img = base;                                        // [slot][hi|lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(x_img + (size_t)NU_SLOTS * SLOT_BYTES);
    uint64_t* slot_full = bars;        // [3] producers -> MMA
    uint64_t* slot_free = bars + 3;    // [3] MMA (commit) -> producers
    uint64_t* d_full = bars + 6;       // MMA -> epilogue: accumulator of a Linear complete (phases alternate first / second Linear)
    uint64_t* h_ready = bars + 7;      // epilogue -> MMA: Swish(y1_pre) image written, accumulator drained
    uint64_t* acc_free = bars + 8;     // epilogue -> MMA: accumulator of the second Linear drained
    uint64_t* w_full = bars + 9;       // epilogue warps -> MMA: weights in tensor memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.rows, 128);
    const int nt = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tiles of this CTA (>= 1)

    if (tid == 0) {
        for (int q = 0; q < NU_SLOTS; ++q) {
            umma::mbar_init(&slot_full[q], NU_PROD_WARPS * 32);
            umma::mbar_init(&slot_free[q], 1);
        }
        umma::mbar_init(d_full, 1);
        umma::mbar_init(h_ready, NU_EPI_WARPS * 32);
        umma::mbar_init(acc_free, NU_EPI_WARPS * 32);
        umma::mbar_init(w_full, NU_EPI_WARPS * 32);
        umma::fence_barrier_init();
    }
    if (warp == NU_MMA_WARP) umma::tmem_alloc(tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < NU_EPI_WARPS) {
        umma::reg_inc<112>();
        // =========================== epilogue: thread = output channel n; warps 0-3 rows 0-63, warps 4-7 rows 64-127 ====
        const int n = tid & 127, hf = warp >> 2;
        const uint32_t lane_q = (uint32_t)((warp & 3) * 32) << 16;
        // ---- weights -> tensor memory, once: this warp's lane quadrant, hi (warps 0-3) or lo (warps 4-7) halves of the three tiles
        if (hf < NSPLIT) {
#pragma unroll 1
            for (int w = 0; w < 3; ++w) {
                const uint4* src = reinterpret_cast<const uint4*>(a.wimg) + (size_t)w * 4096 + (size_t)hf * 2048 + n;
#pragma unroll 1
                for (int pc = 0; pc < 2; ++pc) {          // K pairs 0-31 | 32-63
                    uint4 v[8];
#pragma unroll
                    for (int w4 = 0; w4 < 8; ++w4) v[w4] = src[(pc * 8 + w4) * 128];
                    float f[32];
#pragma unroll
                    for (int w4 = 0; w4 < 8; ++w4) {
                        f[w4 * 4 + 0] = __uint_as_float(v[w4].x); f[w4 * 4 + 1] = __uint_as_float(v[w4].y);
                        f[w4 * 4 + 2] = __uint_as_float(v[w4].z); f[w4 * 4 + 3] = __uint_as_float(v[w4].w);
                    }
                    umma::tmem_st32(tmem + lane_q + NU_TMEM_W + (uint32_t)(w * 64 * NSPLIT + hf * 64 + pc * 32), f);
                }
            }
        }
        umma::tc_fence_before();
        umma::mbar_arrive(w_full);
        const float b3 = a.b3[n], b4 = a.b4[n];
        float wv[NU_MAXV];
#pragma unroll
        for (int t = 0; t < NU_MAXV; ++t) wv[t] = t < a.nv ? a.w3tail[(int64_t)n * a.wt_sn + (int64_t)t * a.wt_st] : 0.f;
        const int nv = a.nv;
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 128;
            const int nr = (int)((a.rows - r0) < 128 ? (a.rows - r0) : 128);
            const int rlim = nr - hf * 64;                   // valid rows among the 64 of this half
            const int64_t rb = r0 + hf * 64;
            const int sa = (2 * it + 1) % NU_SLOTS;          // slot of the agg tile = slot of the Swish(y1_pre) image
            unsigned char* hrow = x_img + (size_t)sa * SLOT_BYTES + n * 128;
            const uint32_t tacc = tmem + lane_q + (uint32_t)(hf * 64);
            // ---------------- first Linear: y1_pre, Swish -> operand image of the second ----------------
            {
                float* yp = a.y1_pre + rb * 128 + n;
                // var of this half's 64 rows: lane l holds rows l and l + 32 (coalesced, requested before the accumulator is
                // waited for); row i's value reaches every lane by shuffle
                float vr[2][NU_MAXV];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
                    for (int t = 0; t < NU_MAXV; ++t) {
                        const int row = h2 * 32 + lane;
                        vr[h2][t] = (t < nv && row < rlim) ? a.var[(rb + row) * nv + t] : 0.f;
                    }
                }
                umma::mbar_wait(d_full, 0);
                umma::tc_fence_after();
                if (warp == 0) NUTL(2, it, 0);
#pragma unroll 1
                for (int c = 0; c < 64; c += 16) {
                    float v[16];
                    umma::tmem_ld16(tacc + (uint32_t)c, v);
                    // rows past the end of the problem carry the bias only; their image rows are never stored as results
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += b3;
                    if (nv > 0) {
#pragma unroll
                        for (int t = 0; t < NU_MAXV; ++t) {
                            if (t < nv) {
                                const float src = c < 32 ? vr[0][t] : vr[1][t];
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = fmaf(__shfl_sync(0xffffffffu, src, (c & 31) + i), wv[t], v[i]);
                            }
                        }
                    }
                    const int lim = rlim - c;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < lim) yp[(c + i) * 128] = v[i];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = swish_tc<FAST>(v[i]);
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        // h[row = cg .. cg+7][k = n] as an MN-major image: row n of the image, 16 bytes
                        const int cg = hf * 64 + c + g * 8;
                        const uint32_t off = (uint32_t)(cg >> 6) * (128u * 128u) + (uint32_t)((((cg & 63) >> 3) ^ (n & 7)) << 4);
                        uint4 hi, lo;
                        if (NSPLIT == 1) {
                            hi.x = umma::pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
                            hi.y = umma::pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
                            hi.z = umma::pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
                            hi.w = umma::pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
                            *reinterpret_cast<uint4*>(hrow + off) = hi;
                        } else {
                            split2_bf16(v[g * 8 + 0], v[g * 8 + 1], hi.x, lo.x);
                            split2_bf16(v[g * 8 + 2], v[g * 8 + 3], hi.y, lo.y);
                            split2_bf16(v[g * 8 + 4], v[g * 8 + 5], hi.z, lo.z);
                            split2_bf16(v[g * 8 + 6], v[g * 8 + 7], hi.w, lo.w);
                            *reinterpret_cast<uint4*>(hrow + off) = hi;
                            *reinterpret_cast<uint4*>(hrow + TILE_BYTES + off) = lo;
                        }
                    }
                }
                umma::fence_async_smem();
                umma::tc_fence_before();
                umma::mbar_arrive(h_ready);
                if (warp == 0) NUTL(2, it, 1);
            }
            // ---------------- second Linear: y2_pre, out = x + Swish(y2_pre) ----------------
            {
                float* yp = a.y2_pre + rb * 128 + n;
                float* yo = a.out + rb * 128 + n;
                const float* xp = a.x + rb * 128 + n;
                float res[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) res[i] = i < rlim ? xp[i * 128] : 0.f;      // in flight while the MMAs run
                umma::mbar_wait(d_full, 1);
                umma::tc_fence_after();
                if (warp == 0) NUTL(2, it, 2);
                // the whole accumulator part of this thread goes to registers first and the accumulator back to the MMA warp:
                // the first Linear of the next tile runs under this tile's stores (HBM-bound: 128 KB out, 64 KB in per tile)
                float v[64];
                {
                    float t0[16], t1[16], t2[16], t3[16];
                    umma::tmem_ld16(tacc, t0);
                    umma::tmem_ld16(tacc + 16u, t1);
                    umma::tmem_ld16(tacc + 32u, t2);
                    umma::tmem_ld16(tacc + 48u, t3);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { v[i] = t0[i]; v[16 + i] = t1[i]; v[32 + i] = t2[i]; v[48 + i] = t3[i]; }
                }
                umma::tc_fence_before();
                umma::mbar_arrive(acc_free);
#pragma unroll
                for (int c = 0; c < 64; c += 16) {
                    const int lim = rlim - c;
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[c + i] += b4;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < lim) yp[(c + i) * 128] = v[c + i];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[c + i] = swish_tc<FAST>(v[c + i]) + res[i];
                    if (c + 16 < 64) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) res[i] = c + 16 + i < rlim ? xp[(c + 16 + i) * 128] : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < lim) yo[(c + i) * 128] = v[c + i];
                }
                if (warp == 0) NUTL(2, it, 3);
            }
        }
    } else if (warp < NU_PROD_WARP0 && warp != NU_MMA_WARP) {
        umma::reg_dec<40>();          // padding warps of the MMA warpgroup
    } else if (warp == NU_MMA_WARP) {
        umma::reg_dec<40>();
        // =========================== MMA issue =======================================================
        const uint32_t id_k = umma::idesc_bf16(128, 128, 0, 0);       // first Linear: B K-major (rows of x / agg)
        const uint32_t id_m = umma::idesc_bf16(128, 128, 0, 1);       // second Linear: B MN-major ([k][row] image)
        const uint64_t xk_d = umma::desc_sw128(umma::smem_u32(x_img), 16, 1024);
        const uint64_t xm_d = umma::desc_sw128(umma::smem_u32(x_img), 128 * 128, 1024);
        constexpr uint32_t TB = TILE_BYTES >> 4, SB = SLOT_BYTES >> 4;
        const uint32_t d = tmem;
        umma::mbar_wait(w_full, 0);
        umma::tc_fence_after();
        // one Linear's K-tile: D (+)= W_tile X;  hi*hi, hi*lo, lo*hi
        auto issue = [&](uint32_t w_t, uint64_t xd, bool k_major, bool first) {
#pragma unroll
            for (int term = 0; term < NTERM; ++term) {
                const uint32_t wa = w_t + (term == 2 ? 64u : 0u);
                const uint64_t xb = xd + (term == 1 ? TB : 0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t koff_k = (uint32_t)((k >> 2) * (128 * 128 >> 4) + (k & 3) * 2), koff_m = (uint32_t)(k * 128);
                    umma::mma_bf16_ts(d, wa + (uint32_t)(k * 8), xb + (uint64_t)(k_major ? koff_k : koff_m), k_major ? id_k : id_m,
                                      (first && term == 0 && k == 0) ? 0u : 1u);
                }
            }
        };
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int qx = 2 * it, qa = 2 * it + 1;
            const int sx = qx % NU_SLOTS, sa = qa % NU_SLOTS;
            umma::mbar_wait(&slot_full[sx], (qx / NU_SLOTS) & 1);
            umma::mbar_wait(&acc_free[0], (it & 1) ^ 1);          // accumulator drained by the previous tile's second epilogue
            umma::tc_fence_after();
            NUTL(1, it, 0);
            if (umma::elect_one()) {
                issue(tmem + NU_TMEM_W, xk_d + (uint64_t)((uint32_t)sx * SB), true, true);
                umma::mma_commit(&slot_free[sx]);
            }
            __syncwarp();
            umma::mbar_wait(&slot_full[sa], (qa / NU_SLOTS) & 1);
            umma::tc_fence_after();
            if (umma::elect_one()) {
                issue(tmem + NU_TMEM_W + 64u * NSPLIT, xk_d + (uint64_t)((uint32_t)sa * SB), true, false);
                umma::mma_commit(d_full);
            }
            __syncwarp();
            NUTL(1, it, 1);
            umma::mbar_wait(h_ready, it & 1);
            umma::tc_fence_after();
            NUTL(1, it, 2);
            if (umma::elect_one()) {
                issue(tmem + NU_TMEM_W + 128u * NSPLIT, xm_d + (uint64_t)((uint32_t)sa * SB), false, true);
                umma::mma_commit(&slot_free[sa]);
                umma::mma_commit(d_full);
            }
            __syncwarp();
            NUTL(1, it, 3);
        }
    } else {
        umma::reg_inc<104>();
        // =========================== producers: 16 rows per warp; x tile, then agg tile of every row tile =============
        const int pw = warp - NU_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
#pragma unroll 1
        for (int q = 0; q < 2 * nt; ++q) {
            const int it = q >> 1, slot = q % NU_SLOTS;
            const float* src = (q & 1) ? a.agg : a.x;
            const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 128 + pw * 16;
            const int64_t left = a.rows - r0;
            const int nv = left >= 16 ? 16 : (left > 0 ? (int)left : 0);       // valid rows of this warp
            {   // L2 prefetch of this warp's rows of the CTA's next tile (same source)
                const int64_t rn = r0 + (int64_t)gridDim.x * 128;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int line = j * 32 + lane;
                    const int64_t row = rn + (line >> 2);
                    if (row < a.rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + row * 128 + (line & 3) * 32));
                }
            }
            float4 x[16];
            {
                const float* p = src + (nv > 0 ? r0 : a.rows - 1) * 128 + lane * 4;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    x[r] = *reinterpret_cast<const float4*>(p);
                    if (r + 1 < nv) p += 128;
                }
            }
            if (nv < 16) {        // the last tile of the problem only
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    if (r >= nv) x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint4 hl[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (NSPLIT == 1) {
                    hl[r].x = umma::pack_bf16(x[r].x, x[r].y);
                    hl[r].y = umma::pack_bf16(x[r].z, x[r].w);
                } else {
                    split2_bf16(x[r].x, x[r].y, hl[r].x, hl[r].z);
                    split2_bf16(x[r].z, x[r].w, hl[r].y, hl[r].w);
                }
            }
            if (pw == 0) NUTL(0, it, (q & 1) * 2);
            umma::mbar_wait(&slot_free[slot], ((q / NU_SLOTS) & 1) ^ 1);
            unsigned char* img = x_img + (size_t)slot * SLOT_BYTES;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                *reinterpret_cast<uint2*>(img + off) = make_uint2(hl[r].x, hl[r].y);
                if (NSPLIT == 2) *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = make_uint2(hl[r].z, hl[r].w);
            }
            umma::fence_async_smem();
            umma::mbar_arrive(&slot_full[slot]);
            if (pw == 0) NUTL(0, it, (q & 1) * 2 + 1);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == NU_MMA_WARP) umma::tmem_dealloc(tmem, 512);
}

int launch_node_update_tc(int precision, const NodeUpdateArgs& a, cudaStream_t s) {
    MGB_REQUIRE(a.nv >= 0 && a.nv <= NU_MAXV, "node_update: at most %d var columns", NU_MAXV);
    MGB_REQUIRE(a.rows < ((int64_t)1 << 31), "node_update: row count out of range");
    MGB_REQUIRE(((uintptr_t)a.x % 16) == 0 && ((uintptr_t)a.agg % 16) == 0 && ((uintptr_t)a.wimg % 16) == 0,
                "node_update: x / agg / packed weights must be 16-byte aligned");
    MGB_REQUIRE(a.nv == 0 || a.var != nullptr, "node_update: var is NULL but nv = %d", a.nv);
    if (a.rows <= 0) return MGB_OK;
    const int64_t tiles = ceil_div<int64_t>(a.rows, 128);
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    ProfScope prof(PROF_NODE_GEMM, s);
    if (precision == 2) {
        MGB_CUDA(cudaFuncSetAttribute(node_update_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)node_update_smem<1>()));
        node_update_tc_kernel<1, true><<<grid, NU_THREADS, node_update_smem<1>(), s>>>(a);
    } else {
        MGB_CUDA(cudaFuncSetAttribute(node_update_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)node_update_smem<2>()));
        node_update_tc_kernel<2, false><<<grid, NU_THREADS, node_update_smem<2>(), s>>>(a);
    }
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
