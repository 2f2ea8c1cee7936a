// magnet_b200 — backward of MAgNet's InteractionNetwork edge function on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Reference: autograd through InteractionNetwork.message + aggregate (models/magnet_gnn.py:70-90, edge_fn :61-68, mean :54):
//     z0 = P[i] + Q[j] + We e          h0 = relu(z0)
//     z_l = W_l h_{l-1} + b_l          h_l = relu(z_l)            l = 1..3
//     y  = W4 h3 + b4                  m = LayerNorm(y)           agg_i = mean_{e -> i} m_e
// Given dagg [N,128] the kernels produce dP (summed by destination), dz0 [E,128] (for dQ, dWe, de), dW1..dW4, db1..db4,
// dgamma, dbeta.  Nothing is saved by the forward pass: every tile of 128 edges is RECOMPUTED on the tensor cores, the
// data-gradient chain runs in place in shared memory, and the weight gradients dW_l = sum_e dz_l (x) h_{l-1} accumulate in
// tensor memory across the tiles of a persistent CTA (drained every IB_DRAIN tiles with round-to-nearest adds: the tensor
// core's own accumulation truncates).
//
// Tensor memory holds 512 columns = 4 accumulators of 128: one working accumulator, one that starts as layer 0's
// accumulator (initialised with P + Q by the producers) and then keeps an activation tile as the A operand (TS form) of a
// weight-gradient MMA, and TWO weight-gradient accumulators.  The four weight gradients therefore take two passes over
// the edges, each with its own minimal recompute:
//   pass A (upper layers)  recompute h0..h3, y;  LayerNorm forward + backward -> dy;  dW4 += h3 (x) dy;  dz3 = W4^T dy . [z3>0];
//                          dW3 += h2 (x) dz3;  dz2 = W3^T dz3 . [z2>0]  -> HBM (fp32, aggregation order, the only [E,128]
//                          intermediate);  db4, db3, db2, dgamma, dbeta in registers
//   pass B (lower layers)  recompute h0, h1;  dW2 += h1 (x) dz2;  dz1 = W2^T dz2 . [z1>0];  dW1 += h0 (x) dz1;
//                          dz0 = W1^T dz1 . [z0>0]  -> HBM in COO order (times e_scale) + segmented sum by destination -> dP
// Shared memory: one layer's weight images (64 KB, the forward images serve the transposed products as MN-major
// operands), the working activation / gradient tile (64 KB, rewritten in place) and ONE retained activation tile (64 KB).
// Operands are split into two fp16 values (three MMA terms) as in the forward kernel, which makes the recompute
// bit-identical to the forward pass (same ReLU masks); gradients are brought into the fp16 range by ONE power-of-two
// factor per launch (2^-ceil(log2 max |dagg| / deg)): LayerNorm backward, the ReLU masks and the contractions are linear
// in it, and it is divided out exactly when results leave the SM.
#include "internal.cuh"
#include "tc_common.cuh"
#include "segmeta.cuh"
#include <stdlib.h>

namespace mgb {

constexpr int IB_TE = 128;                       // edge positions per tile
constexpr int IB_EPI_WARPS = 8, IB_PROD_WARPS = 8;
// whole warpgroups, so that registers can follow the work (setmaxnreg): warps 0-7 epilogue, 8 MMA, 9 metadata, 10-11 idle,
// 12-19 producers; 640 threads x 96 registers at launch -> 120 epilogue / 40 MMA warpgroup / 96 producers
constexpr int IB_MMA_WARP = IB_EPI_WARPS, IB_META_WARP = IB_EPI_WARPS + 1, IB_PROD_WARP0 = IB_EPI_WARPS + 4;
constexpr int IB_THREADS = (IB_PROD_WARP0 + IB_PROD_WARPS) * 32;      // 640
constexpr int IB_FLUSH = 64;                     // positions per epilogue warp = granularity of the stored partial sums
constexpr int IB_DRAIN = 8;                      // tiles between drains of the weight-gradient accumulators
constexpr int IB_NVEC_A = 5, IB_NVEC_B = 1;      // per-channel vector gradients of pass A (db4, db3, db2, dgamma, dbeta) / B (db1)
using IbMeta = TileMetaT<IB_TE>;
constexpr size_t IN_EDGE_BWD_SMEM = 1024 + (size_t)6 * TILE_BYTES + 2 * sizeof(IbMeta) + 2 * IB_TE * sizeof(int) +
                                    IB_TE * sizeof(float4) + IB_TE * sizeof(int2) + 128 * sizeof(float) + 256;

struct InEdgeBwdArgs {
    const float* e;            // [E][128] edge features, COO order
    float e_scale;             // 2^l
    const int32_t* perm;       // [E] COO edge id of every aggregation-order position (NULL: identity)
    const float* pq;           // [N][256]  P | Q
    const int32_t* rowptr;
    const int32_t* dstv;
    const int32_t* srcv;
    int64_t n_edges;
    const void* wimg;          // [5][hi | lo] images of We, W1..W4 (the forward pass's)
    const float* bias;         // [5][128]
    const float* gamma;
    const float* beta;
    const float* dagg;         // [N][128] gradient of the aggregated messages
    const uint32_t* gmax_bits; // device scalar: bits of max_i |dagg[i]| / deg(i)
    float* dz2;                // [tiles * 128][128] scaled, aggregation order (pass A out, pass B in)
    float* dz0;                // [E][128] COO order, times e_scale (pass B out)
    float* dpq;                // [N][256]: dP = columns 0..127 (pass B, through the segment metadata)
    float* part_head;
    float* part_tail;
    float* wpart;              // [grid][2][128][128] partial weight gradients of this pass (zeroed by the launcher)
    float* vpart;              // [grid][2 halves][NVEC][128] partial per-channel gradients of this pass
    int* range_flag;
    int dbg;                   // developer build only (-DMGB_IB_DEBUG; env MGB_IB_DEBUG = 1 / 2 / 3: pass A stores y / dy / dz3 instead of dz2)
};

__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// power-of-two factor that brings the gradients of this launch into the fp16 range: max |dm| * s in (0.5, 1]
__device__ __forceinline__ float ib_scale(const uint32_t* gmax_bits) {
    const float g = __uint_as_float(*gmax_bits);
    if (!(g > 0.f) || !(g < 3.0e38f)) return 1.0f;
    int ex;
    frexpf(g, &ex);                       // g = f * 2^ex, f in [0.5, 1)
    ex = ex > 100 ? 100 : (ex < -100 ? -100 : ex);
    return __int_as_float((127 - ex) << 23);     // 2^-ex, exact
}

// 32 values of channel n at positions [c0, c0+32) -> operand image row n (MN-major image [n][e], in place), optionally also
// into tensor memory as the packed A operand of a TS-form MMA (lane = n, 32-bit column j = positions (2j, 2j+1))
template <int NSPLIT>
__device__ __forceinline__ void ib_store32(unsigned char* xrow, int c0, int n, const float (&v)[32], uint32_t ts_addr /*0: none*/) {
    uint32_t th[16], tl[16];
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
            tl[g * 4 + 0] = lo.x; tl[g * 4 + 1] = lo.y; tl[g * 4 + 2] = lo.z; tl[g * 4 + 3] = lo.w;
        } else {
            hi.x = umma::pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
            hi.y = umma::pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
            hi.z = umma::pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
            hi.w = umma::pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
            *reinterpret_cast<uint4*>(xrow + off) = hi;
        }
        th[g * 4 + 0] = hi.x; th[g * 4 + 1] = hi.y; th[g * 4 + 2] = hi.z; th[g * 4 + 3] = hi.w;
    }
    if (ts_addr) {
        tmem_st16u(ts_addr + (uint32_t)(c0 >> 1), th);
        if (NSPLIT == 2) tmem_st16u(ts_addr + 64u + (uint32_t)(c0 >> 1), tl);
    }
}

// hidden-layer epilogue of the recompute: h = relu(D + bias) for the 64 positions of this thread -> image (slot) [+ TS copy],
// returns the ReLU mask of the 64 positions
template <int NSPLIT>
__device__ __forceinline__ uint2 ib_relu_epilogue(uint32_t tacc, float bias, unsigned char* slot, int hf, int n, uint32_t ts_base,
                                                  int* range_flag, bool image) {
    unsigned char* xrow = slot + n * 128;
    uint32_t mask[2];
    float vmax = 0.f;
#pragma unroll 1
    for (int cb = 0; cb < 64; cb += 32) {
        float v[32];
        umma::tmem_ld32(tacc + (uint32_t)cb, v);
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float z = v[i] + bias;
            m |= (z > 0.f ? 1u : 0u) << i;
            v[i] = fmaxf(z, 0.f);
        }
        mask[cb >> 5] = m;
        if (NSPLIT == 2) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) vmax = fmaxf(vmax, fmaxf(v[i], v[i + 1]));
        }
        if (image) {
            ib_store32<NSPLIT>(xrow, hf * 64 + cb, n, v, ts_base);
        } else {                      // tensor-memory copy only
            uint32_t th[16], tl[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (NSPLIT == 2) split2_f16(v[2 * j], v[2 * j + 1], th[j], tl[j]);
                else th[j] = umma::pack_bf16(v[2 * j], v[2 * j + 1]);
            }
            tmem_st16u(ts_base + (uint32_t)((hf * 64 + cb) >> 1), th);
            if (NSPLIT == 2) tmem_st16u(ts_base + 64u + (uint32_t)((hf * 64 + cb) >> 1), tl);
        }
    }
    if (ts_base) tmem_wait_st();
    if (NSPLIT == 2 && vmax >= 32768.f && range_flag) *range_flag = 1;
    return make_uint2(mask[0], mask[1]);
}

// data-gradient epilogue: dz = D . mask for the 64 positions of this thread -> image (slot), returns the sum over the positions
template <int NSPLIT>
__device__ __forceinline__ float ib_mask_epilogue(uint32_t tacc, uint2 mask, unsigned char* slot, int hf, int n, int* range_flag) {
    unsigned char* xrow = slot + n * 128;
    float sum = 0.f, vmax = 0.f;
#pragma unroll 1
    for (int cb = 0; cb < 64; cb += 32) {
        float v[32];
        umma::tmem_ld32(tacc + (uint32_t)cb, v);
        const uint32_t m = cb ? mask.y : mask.x;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            v[i] = (m >> i) & 1u ? v[i] : 0.f;
            sum += v[i];
        }
        if (NSPLIT == 2) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) vmax = fmaxf(vmax, fmaxf(fabsf(v[i]), fabsf(v[i + 1])));
        }
        ib_store32<NSPLIT>(xrow, hf * 64 + cb, n, v, 0u);
    }
    if (NSPLIT == 2 && vmax >= 32768.f && range_flag) *range_flag = 1;
    return sum;
}

// one MMA group D (+)= A B over K = 128 (eight K-steps), three terms for the hi/lo split (small terms first)
//   a_kind: 0 shared K-major, 1 shared MN-major, 2 tensor memory (a = TMEM column address of the hi part, lo at +64)
template <int NSPLIT>
__device__ __forceinline__ void ib_mma_group(uint32_t d, int a_kind, uint64_t a, int b_mn, uint64_t b, uint32_t idesc, uint32_t acc0) {
    constexpr uint32_t TB = TILE_BYTES >> 4;
#pragma unroll
    for (int term = 0; term < (NSPLIT == 2 ? 3 : 1); ++term) {
        const bool a_lo = NSPLIT == 2 && term == 0, b_lo = NSPLIT == 2 && term == 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t koff_k = (uint32_t)((k >> 2) * (128 * 128 >> 4) + (k & 3) * 2), koff_m = (uint32_t)(k * 128);
            const uint64_t bb = b + (b_lo ? TB : 0) + (uint64_t)(b_mn ? koff_m : koff_k);
            const uint32_t acc = (term | k) ? 1u : acc0;
            if (a_kind == 2)
                umma::mma_bf16_ts(d, (uint32_t)a + (a_lo ? 64u : 0u) + (uint32_t)(k * 8), bb, idesc, acc);
            else
                umma::mma_bf16(d, a + (a_lo ? TB : 0) + (uint64_t)(a_kind == 1 ? koff_m : koff_k), bb, idesc, acc);
        }
    }
}

// The intermediate-quantity dumps of tools/dev_in_bwd_dbg.py exist in the developer build only: as a run-time switch their
// per-position tests were 5 % of the instructions of pass A (ncu source view).
#ifdef MGB_IB_DEBUG
#define IB_DBG(a) ((a).dbg)
#else
#define IB_DBG(a) 0
#endif

#ifdef MGB_TIMELINE
__device__ long long* g_ib_timeline = nullptr;      // [pass 0..1][tile 0..3][role 0..2][event 0..31] clock64 of CTA 0
#define IBTL(role, it_, ev) do { if (blockIdx.x == 0 && (it_) < 4 && (threadIdx.x & 31) == 0 && g_ib_timeline && (ev) < 32) g_ib_timeline[(((PASS) * 4 + (it_)) * 3 + (role)) * 32 + (ev)] = clock64(); } while (0)
int set_ib_timeline_buffer(long long* p) {
    return cudaMemcpyToSymbol(g_ib_timeline, &p, sizeof(p)) == cudaSuccess ? MGB_OK : MGB_ERR_CUDA;
}
#else
#define IBTL(role, it_, ev) do { } while (0)
#endif

template <int PASS /*0 = A (upper layers), 1 = B (lower layers)*/, int NSPLIT>
__global__ void __launch_bounds__(IB_THREADS, 1) in_edge_bwd_tc_kernel(const InEdgeBwdArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* w_img = base;                                        // [hi|lo] of the current layer
    unsigned char* x_img = base + (size_t)2 * TILE_BYTES;               // working tile [hi|lo]; fp32 staging for LayerNorm
    unsigned char* h_img = base + (size_t)4 * TILE_BYTES;               // retained activation tile [hi|lo]
    IbMeta* metas = reinterpret_cast<IbMeta*>(base + (size_t)6 * TILE_BYTES);
    int* rowid = reinterpret_cast<int*>(metas + 2);                     // [2][128] COO row of every position (-1: none)
    float4* stats = reinterpret_cast<float4*>(rowid + 2 * IB_TE);       // [128] mean, rstd, c1, c2 of every edge
    int2* dinv = reinterpret_cast<int2*>(stats + IB_TE);                // [128] destination node, bits of scale / in-degree
    float* gam_s = reinterpret_cast<float*>(dinv + IB_TE);              // [128] LayerNorm gain
    uint64_t* bars = reinterpret_cast<uint64_t*>(gam_s + 128);
    uint64_t* x_full = bars;          // producers -> MMA: layer-0 operand written, accumulator initialised
    uint64_t* x_free = bars + 1;      // MMA -> producers: the tile's last MMAs are done (working tile + TS region free)
    uint64_t* t_full = bars + 2;      // MMA -> epilogue
    uint64_t* x_ready = bars + 3;     // epilogue -> MMA
    uint64_t* w_bar = bars + 4;
    uint64_t* w_free = bars + 5;
    uint64_t* x_full2 = bars + 6;     // pass B: dz2 tile written
    uint64_t* x_free2 = bars + 7;     // pass B: layer 0's MMAs are done with the e tile
    uint64_t* m_full = bars + 8;      // [2] meta warp -> epilogue
    uint64_t* m_empty = bars + 10;    // [2]
    uint64_t* t_full0 = bars + 12;    // MMA -> epilogue, layer 0 only: it may complete while the previous tile's last epilogue still
                                      // runs, and a barrier must never get two phases ahead of its waiter
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_edges, IB_TE);
    const int nt = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tiles of this CTA (>= 1)

    if (tid == 0) {
        umma::mbar_init(x_full, IB_PROD_WARPS * 32);
        umma::mbar_init(x_free, 1);
        umma::mbar_init(t_full, 1);
        umma::mbar_init(t_full0, 1);
        umma::mbar_init(x_ready, IB_EPI_WARPS * 32);
        umma::mbar_init(w_bar, 1);
        umma::mbar_init(w_free, 1);
        umma::mbar_init(x_full2, IB_PROD_WARPS * 32);
        umma::mbar_init(x_free2, 1);
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&m_full[s], 32);
            umma::mbar_init(&m_empty[s], IB_EPI_WARPS * 32);
        }
        umma::fence_barrier_init();
    }
    if (tid < 128) gam_s[tid] = a.gamma[tid];
    if (warp == IB_MMA_WARP) umma::tmem_alloc(tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // tensor-memory columns: [0,128) working accumulator R1; [128,256) R2 = layer-0 accumulator, later the packed TS operand;
    // [256,384) and [384,512) the two weight-gradient accumulators of the pass (transposed: lane = input k, column = output n)
    const float gs = ib_scale(a.gmax_bits), inv_gs = 1.0f / gs;

    if (warp < IB_EPI_WARPS) {
        umma::reg_inc<120>();        // (120 - 96) x 256 <= (96 - 40) x 128 released by the MMA warpgroup
        // =========================== epilogue: thread = channel n; warps 0-3 positions 0-63, warps 4-7 positions 64-127 ====
        const int n = tid & 127, hf = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t R1 = tmem + lane_base + (uint32_t)(hf * 64), R2 = tmem + 128u + lane_base + (uint32_t)(hf * 64);
        const uint32_t TS = tmem + 128u + lane_base;                    // packed operand: columns [0,64) hi, [64,128) lo
        const float b1 = a.bias[128 + n], b2 = a.bias[256 + n], b3 = a.bias[384 + n], b4 = a.bias[512 + n];
        const float gamma = a.gamma[n];
        float acc_v[PASS == 0 ? IB_NVEC_A : IB_NVEC_B];
#pragma unroll
        for (int i = 0; i < (PASS == 0 ? IB_NVEC_A : IB_NVEC_B); ++i) acc_v[i] = 0.f;
        uint32_t tf = 0;                 // completed phases of t_full
        int tl_it = 0, tl_ev = 0;        // (developer timeline: event counter of the current tile)
        auto wait_t = [&]() {
            umma::mbar_wait(t_full, tf & 1);
            ++tf;
            umma::tc_fence_after();
            if (warp == 0) IBTL(0, tl_it, tl_ev++);
        };
        auto wait_t0 = [&](int it_) {
            tl_it = it_; tl_ev = 0;
            if (warp == 0) IBTL(0, tl_it, tl_ev++);
            umma::mbar_wait(t_full0, it_ & 1);
            umma::tc_fence_after();
            if (warp == 0) IBTL(0, tl_it, tl_ev++);
        };
        auto signal = [&]() {
            umma::fence_async_smem();
            umma::tc_fence_before();
            if (warp == 0) IBTL(0, tl_it, tl_ev++);
            umma::mbar_arrive(x_ready);
        };
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            auto dump = [&](uint32_t tacc, float bias) {          // developer switch: h_l of pass A instead of dz2
                float v[32];
                for (int cb = 0; cb < 64; cb += 32) {
                    umma::tmem_ld32(tacc + (uint32_t)cb, v);
                    for (int i = 0; i < 32; ++i) a.dz2[(tile * IB_TE + hf * 64 + cb + i) * 128 + n] = fmaxf(v[i] + bias, 0.f);
                }
            };
            if constexpr (PASS == 0) {
                // ---- recompute: h0 (accumulator R2, initialised with P + Q), h1 -> working tile; h2 -> retained tile;
                // h3 -> working tile + tensor memory (A operand of dW4)
                wait_t0(it);
                if (IB_DBG(a) == 4) dump(R2, 0.f);
                ib_relu_epilogue<NSPLIT>(R2, 0.f, x_img, hf, n, 0u, a.range_flag, true);
                signal();
                wait_t();
                if (IB_DBG(a) == 5) dump(R1, b1);
                ib_relu_epilogue<NSPLIT>(R1, b1, x_img, hf, n, 0u, a.range_flag, true);
                signal();
                wait_t();
                if (IB_DBG(a) == 6) dump(R1, b2);
                const uint2 mask2 = ib_relu_epilogue<NSPLIT>(R1, b2, h_img, hf, n, 0u, a.range_flag, true);
                signal();
                wait_t();
                if (IB_DBG(a) == 7) dump(R1, b3);
                const uint2 mask3 = ib_relu_epilogue<NSPLIT>(R1, b3, x_img, hf, n, TS, a.range_flag, true);
                signal();
                // ---- y = D + b4: LayerNorm forward statistics and backward, dy -> working tile
                wait_t();
                if (tid < IB_TE) {
                    const int64_t p = tile * IB_TE + tid;
                    int d = 0;
                    float inv = 0.f;
                    if (p < a.n_edges) {
                        d = a.dstv[p];
                        inv = gs / (float)(a.rowptr[d + 1] - a.rowptr[d]);
                    }
                    dinv[tid] = make_int2(d, __float_as_int(inv));
                }
#pragma unroll 1
                for (int cb = 0; cb < 64; cb += 16) {          // pass 1: y[n][e] -> fp32 staging [e][n] in the working tile
                    const int c0 = hf * 64 + cb;
                    float v[16];
                    umma::tmem_ld16(R1 + (uint32_t)cb, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int e = c0 + i;
                        *reinterpret_cast<float*>(x_img + e * 512 + ((((n >> 2) ^ (e & 31))) << 4) + (n & 3) * 4) = v[i] + b4;
                        if (IB_DBG(a) == 1) a.dz2[(tile * IB_TE + e) * 128 + n] = v[i] + b4;
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                {                                               // pass 2: two threads per edge (64 channels each)
                    const int e = tid >> 1, part = tid & 1;
                    const unsigned char* row = x_img + e * 512;
                    const int2 di = dinv[e];
                    const float* drow = a.dagg + (int64_t)di.x * 128;
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
                    float q = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int chunk = part * 16 + (i ^ (part << 2));
                        const float4 dg = __ldg(reinterpret_cast<const float4*>(drow) + chunk);
                        const float4 gm = *reinterpret_cast<const float4*>(gam_s + chunk * 4);
                        const float dx = y[i].x - mean, dy = y[i].y - mean, dz = y[i].z - mean, dw = y[i].w - mean;
                        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                        const float g0 = dg.x * gm.x, g1 = dg.y * gm.y, g2 = dg.z * gm.z, g3 = dg.w * gm.w;
                        s1 += (g0 + g1) + (g2 + g3);
                        s2 += (g0 * dx + g1 * dy) + (g2 * dz + g3 * dw);
                    }
                    q += __shfl_xor_sync(0xffffffffu, q, 1);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                    const float rstd = 1.0f / sqrtf(q * (1.0f / 128.0f) + 1e-5f);
                    const float inv = __int_as_float(di.y);
                    if (part == 0) stats[e] = make_float4(mean, rstd, s1 * inv * (1.0f / 128.0f), s2 * inv * rstd * (1.0f / 128.0f));
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                {                                               // pass 3: dy from the accumulator, written as the next operand
                    unsigned char* xrow = x_img + n * 128;
                    float vmax = 0.f;
#pragma unroll 1
                    for (int cb = 0; cb < 64; cb += 32) {
                        float v[32];
                        umma::tmem_ld32(R1 + (uint32_t)cb, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int e = hf * 64 + cb + i;
                            const float4 st = stats[e];
                            const int2 di = dinv[e];
                            const float dm = __ldg(a.dagg + (int64_t)di.x * 128 + n) * __int_as_float(di.y);
                            const float xh = ((v[i] + b4) - st.x) * st.y;
                            const float g = dm * gamma;
                            const float dy = st.y * ((g - st.z) - xh * st.w);
                            acc_v[3] = fmaf(dm, xh, acc_v[3]);
                            acc_v[4] += dm;
                            acc_v[0] += dy;
                            v[i] = dy;
                            if (IB_DBG(a) == 2) a.dz2[(tile * IB_TE + e) * 128 + n] = dy;
                        }
                        if (NSPLIT == 2) {
#pragma unroll
                            for (int i = 0; i < 32; i += 2) vmax = fmaxf(vmax, fmaxf(fabsf(v[i]), fabsf(v[i + 1])));
                        }
                        ib_store32<NSPLIT>(xrow, hf * 64 + cb, n, v, 0u);
                    }
                    if (NSPLIT == 2 && vmax >= 32768.f && a.range_flag) *a.range_flag = 1;
                }
                signal();
                // ---- dz3 = (W4^T dy) . [z3 > 0] -> working tile
                wait_t();
                acc_v[1] += ib_mask_epilogue<NSPLIT>(R1, mask3, x_img, hf, n, a.range_flag);
                if (IB_DBG(a) == 3) {
                    float v[32];
                    for (int cb = 0; cb < 64; cb += 32) {
                        umma::tmem_ld32(R1 + (uint32_t)cb, v);
                        for (int i = 0; i < 32; ++i) a.dz2[(tile * IB_TE + hf * 64 + cb + i) * 128 + n] = (((cb ? mask3.y : mask3.x) >> i) & 1u) ? v[i] : 0.f;
                    }
                }
                signal();
                // ---- dz2 = (W3^T dz3) . [z2 > 0] -> HBM (scaled; aggregation order, whole tiles)
                wait_t();
                {
                    float* out = a.dz2 + (tile * IB_TE + hf * 64) * 128 + n;
                    float sum = 0.f;
#pragma unroll 1
                    for (int cb = 0; cb < 64; cb += 32) {
                        float v[32];
                        umma::tmem_ld32(R1 + (uint32_t)cb, v);
                        const uint32_t m = cb ? mask2.y : mask2.x;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float z = (m >> i) & 1u ? v[i] : 0.f;
                            sum += z;
                            if (IB_DBG(a) == 0) out[(cb + i) * 128] = z;
                        }
                    }
                    acc_v[2] += sum;
                }
            } else {
                // ---- recompute: h0 -> retained tile, h1 -> tensor memory (A operand of dW2)
                wait_t0(it);
                const uint2 mask0 = ib_relu_epilogue<NSPLIT>(R2, 0.f, h_img, hf, n, 0u, a.range_flag, true);
                signal();
                wait_t();
                const uint2 mask1 = ib_relu_epilogue<NSPLIT>(R1, b1, x_img, hf, n, TS, a.range_flag, false);
                signal();
                // ---- dz1 = (W2^T dz2) . [z1 > 0] -> working tile
                wait_t();
                acc_v[0] += ib_mask_epilogue<NSPLIT>(R1, mask1, x_img, hf, n, a.range_flag);
                signal();
                // ---- dz0 = (W1^T dz1) . [z0 > 0]: unscaled, times e_scale -> HBM in COO order; segmented sum by destination -> dP
                wait_t();
                const int slot = it & 1;
                umma::mbar_wait(&m_full[slot], (it >> 1) & 1);
                {
                    const IbMeta* M = metas + slot;
                    const int* rid = rowid + slot * IB_TE;
                    float sum = 0.f;
#pragma unroll 1
                    for (int cb = 0; cb < 64; cb += 8) {
                        const int c0 = hf * 64 + cb;
                        float v[8];
                        umma::tmem_ld8(R1 + (uint32_t)cb, v);
                        const uint32_t mb = ((cb & 32) ? mask0.y : mask0.x) >> (cb & 31);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            v[i] = (mb >> i) & 1u ? v[i] * inv_gs : 0.f;
                            const int r = rid[c0 + i];
                            if (r >= 0) a.dz0[(int64_t)r * 128 + n] = v[i] * a.e_scale;
                        }
                        uint32_t fm = (M->flushmask[c0 >> 5] >> (c0 & 31)) & 0xffu, todo = 0xffu;
                        if (fm == 0) {
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
                umma::tc_fence_before();
                umma::mbar_arrive(&m_empty[slot]);
            }
            if (warp == 0) IBTL(0, tl_it, tl_ev++);
            // ---- drain the weight-gradient accumulators into this CTA's partial (round-to-nearest adds, fixed order)
            if (((it + 1) % IB_DRAIN) == 0 || it + 1 == nt) {
#pragma unroll 1
                for (int g = 0; g < 2; ++g) {
                    const uint32_t acc = tmem + 256u + (uint32_t)(g * 128) + lane_base + (uint32_t)(hf * 64);
                    float* wp = a.wpart + ((size_t)blockIdx.x * 2 + g) * 128 * 128 + (size_t)(hf * 64) * 128 + n;
#pragma unroll 1
                    for (int cb = 0; cb < 64; cb += 32) {
                        float v[32];
                        umma::tmem_ld32(acc + (uint32_t)cb, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) wp[(cb + i) * 128] += v[i] * inv_gs;
                    }
                }
                umma::tc_fence_before();
            }
        }
        // per-channel gradients of this CTA (two halves of the positions: two rows)
        constexpr int NV = PASS == 0 ? IB_NVEC_A : IB_NVEC_B;
#pragma unroll
        for (int i = 0; i < NV; ++i) a.vpart[(((size_t)blockIdx.x * 2 + hf) * NV + i) * 128 + n] = acc_v[i] * inv_gs;
    } else if (warp < IB_PROD_WARP0 && warp != IB_MMA_WARP && warp != IB_META_WARP) {
        umma::reg_dec<40>();          // padding warps of the MMA / metadata warpgroup
    } else if (warp == IB_MMA_WARP) {
        umma::reg_dec<40>();
        // =========================== MMA issue + weight loads =======================================
        const uint32_t id_kk = NSPLIT == 2 ? umma::idesc_f16(128, 128, 0, 0) : umma::idesc_bf16(128, 128, 0, 0);   // A K-major,  B K-major
        const uint32_t id_km = NSPLIT == 2 ? umma::idesc_f16(128, 128, 0, 1) : umma::idesc_bf16(128, 128, 0, 1);   // A K-major,  B MN-major
        const uint32_t id_mk = NSPLIT == 2 ? umma::idesc_f16(128, 128, 1, 0) : umma::idesc_bf16(128, 128, 1, 0);   // A MN-major, B K-major
        const uint32_t id_mm = NSPLIT == 2 ? umma::idesc_f16(128, 128, 1, 1) : umma::idesc_bf16(128, 128, 1, 1);   // A MN-major, B MN-major
        const uint64_t w_k = umma::desc_sw128(umma::smem_u32(w_img), 16, 1024), w_m = umma::desc_sw128(umma::smem_u32(w_img), 128 * 128, 1024);
        const uint64_t x_k = umma::desc_sw128(umma::smem_u32(x_img), 16, 1024), x_m = umma::desc_sw128(umma::smem_u32(x_img), 128 * 128, 1024);
        const uint64_t h_k = umma::desc_sw128(umma::smem_u32(h_img), 16, 1024), h_m = umma::desc_sw128(umma::smem_u32(h_img), 128 * 128, 1024);
        const uint32_t R1 = tmem, R2 = tmem + 128u, GA = tmem + 256u, GB = tmem + 384u;
        uint32_t wl = 0, xr = 0;         // weight loads issued, x_ready phases consumed
        int ml_it = 0, ml_ev = 0;
        auto load_w = [&](int layer) {
            if (wl > 0) umma::mbar_wait(w_free, (wl - 1) & 1);          // the MMAs that read the previous images are done
            IBTL(1, ml_it, ml_ev++);                                    // previous group complete
            if (lane == 0) load_w2_image(w_img, (const unsigned char*)a.wimg + (size_t)layer * 2 * TILE_BYTES, NSPLIT * TILE_BYTES, w_bar);
            umma::mbar_wait(w_bar, wl & 1);
            ++wl;
            IBTL(1, ml_it, ml_ev++);                                    // weights landed
        };
        auto wait_x = [&]() {
            umma::mbar_wait(x_ready, xr & 1);
            ++xr;
            umma::tc_fence_after();
            IBTL(1, ml_it, ml_ev++);                                    // operand ready
        };
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const uint32_t g0 = (it % IB_DRAIN) == 0 ? 0u : 1u;        // first tile after a drain: the gradient accumulators restart
            ml_it = it; ml_ev = 0;
            // ---- layer 0: R2 (= P + Q) += We e
            load_w(0);
            umma::mbar_wait(x_full, it & 1);
            umma::tc_fence_after();
            IBTL(1, ml_it, ml_ev++);                                    // x_full
            if (umma::elect_one()) {
                ib_mma_group<NSPLIT>(R2, 0, w_k, 0, x_k, id_kk, 1u);
                umma::mma_commit(t_full0);
                umma::mma_commit(w_free);
                if (PASS == 1) umma::mma_commit(x_free2);
            }
            __syncwarp();
            if constexpr (PASS == 0) {
                // ---- layers 1, 2 read the working tile, layer 3 the retained tile (h2), layer 4 the working tile (h3)
#pragma unroll 1
                for (int l = 1; l <= 4; ++l) {
                    load_w(l);
                    wait_x();
                    if (umma::elect_one()) {
                        ib_mma_group<NSPLIT>(R1, 0, w_k, 1, l == 3 ? h_m : x_m, id_km, 0u);
                        umma::mma_commit(t_full);
                        if (l < 4) umma::mma_commit(w_free);            // W4 stays for the data gradient below
                    }
                    __syncwarp();
                }
                // ---- dW4^T += h3 (tensor memory) x dy;  R1 = W4^T dy
                wait_x();
                if (umma::elect_one()) {
                    ib_mma_group<NSPLIT>(GA, 2, (uint64_t)R2, 0, x_k, id_kk, g0);
                    ib_mma_group<NSPLIT>(R1, 1, w_m, 1, x_m, id_mm, 0u);
                    umma::mma_commit(t_full);
                    umma::mma_commit(w_free);
                }
                __syncwarp();
                // ---- dW3^T += h2 (retained tile) x dz3;  R1 = W3^T dz3
                load_w(3);
                wait_x();
                if (umma::elect_one()) {
                    ib_mma_group<NSPLIT>(GB, 0, h_k, 0, x_k, id_kk, g0);
                    ib_mma_group<NSPLIT>(R1, 1, w_m, 1, x_m, id_mm, 0u);
                    umma::mma_commit(t_full);
                    umma::mma_commit(w_free);
                    umma::mma_commit(x_free);
                }
                __syncwarp();
            } else {
                // ---- layer 1 reads the retained tile (h0)
                load_w(1);
                wait_x();
                if (umma::elect_one()) {
                    ib_mma_group<NSPLIT>(R1, 0, w_k, 1, h_m, id_km, 0u);
                    umma::mma_commit(t_full);
                    umma::mma_commit(w_free);
                }
                __syncwarp();
                // ---- dW2^T += h1 (tensor memory) x dz2 (rows of the working tile: K = e along the rows);  R1 = W2^T dz2
                load_w(2);
                wait_x();
                umma::mbar_wait(x_full2, it & 1);
                umma::tc_fence_after();
                if (umma::elect_one()) {
                    ib_mma_group<NSPLIT>(GA, 2, (uint64_t)R2, 1, x_m, id_km, g0);
                    ib_mma_group<NSPLIT>(R1, 1, w_m, 0, x_k, id_mk, 0u);
                    umma::mma_commit(t_full);
                    umma::mma_commit(w_free);
                }
                __syncwarp();
                // ---- dW1^T += h0 (retained tile) x dz1;  R1 = W1^T dz1
                load_w(1);
                wait_x();
                if (umma::elect_one()) {
                    ib_mma_group<NSPLIT>(GB, 0, h_k, 0, x_k, id_kk, g0);
                    ib_mma_group<NSPLIT>(R1, 1, w_m, 1, x_m, id_mm, 0u);
                    umma::mma_commit(t_full);
                    umma::mma_commit(w_free);
                    umma::mma_commit(x_free);
                }
                __syncwarp();
            }
        }
    } else if (warp == IB_META_WARP) {
        umma::reg_dec<40>();
        // =========================== pass B: segment metadata + COO rows, one tile ahead ===============================
        if (PASS == 1) {
#pragma unroll 1
            for (int it = 0; it < nt; ++it) {
                const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
                const int slot = it & 1;
                umma::mbar_wait_relaxed<2000>(&m_empty[slot], ((it >> 1) & 1) ^ 1);
                build_tile_meta(metas + slot, a.rowptr, a.dstv, a.srcv, a.n_edges, tile, lane, a.dpq, 256, false, a.part_head,
                                a.part_tail, IB_FLUSH);
#pragma unroll
                for (int j = 0; j < IB_TE / 32; ++j) {
                    const int64_t p = tile * IB_TE + j * 32 + lane;
                    rowid[slot * IB_TE + j * 32 + lane] = p < a.n_edges ? (a.perm ? a.perm[p] : (int)p) : -1;
                }
                umma::mbar_arrive(&m_full[slot]);
            }
        }
    } else {
        // =========================== producers: layer-0 operand + accumulator initialisation (+ the dz2 tile in pass B) =====
        const int pw = warp - IB_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        const float sc = a.e_scale;
        const int quad = warp & 3, chalf = pw >> 2;            // TMEM lane quadrant of this warp, its half of the positions
        const int n = quad * 32 + lane;
        const float* pqn = a.pq + n;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        // 16 rows of 128 floats (one float4 per lane) -> fp16 hi | lo K-major image rows pw*16 .. pw*16+15 of the working tile
        auto rows_to_regs = [&](const float* src, int64_t row_of_lane /*lanes 0-15: row, -1 none*/, float scale, uint4 (&hl)[16]) {
            float4 x[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int64_t c = __shfl_sync(0xffffffffu, row_of_lane, r);
                x[r] = *reinterpret_cast<const float4*>(src + (c < 0 ? 0 : c) * 128 + lane * 4);
                if (c < 0) x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                float4 h = x[r];
                h.x *= scale; h.y *= scale; h.z *= scale; h.w *= scale;
                if (NSPLIT == 2) {
                    if (fmaxf(fmaxf(fabsf(h.x), fabsf(h.y)), fmaxf(fabsf(h.z), fabsf(h.w))) >= 32768.f && a.range_flag) *a.range_flag = 1;
                    split2_f16(h.x, h.y, hl[r].x, hl[r].z);
                    split2_f16(h.z, h.w, hl[r].y, hl[r].w);
                } else {
                    hl[r].x = umma::pack_bf16(h.x, h.y);
                    hl[r].y = umma::pack_bf16(h.z, h.w);
                }
            }
        };
        auto regs_to_image = [&](const uint4 (&hl)[16]) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                *reinterpret_cast<uint2*>(x_img + off) = make_uint2(hl[r].x, hl[r].y);
                if (NSPLIT == 2) *reinterpret_cast<uint2*>(x_img + TILE_BYTES + off) = make_uint2(hl[r].z, hl[r].w);
            }
        };
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t p0 = tile * IB_TE + pw * 16;
            int64_t mine = -1;
            if (lane < 16 && p0 + lane < a.n_edges) mine = a.perm ? (int64_t)a.perm[p0 + lane] : p0 + lane;
            uint4 hl[16];
            rows_to_regs(a.e, mine, sc, hl);
            // P | Q rows of this warp's 64 positions are pulled towards the SM while the slot is still busy
            if (quad == 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t p = tile * IB_TE + chalf * 64 + h * 32 + lane;
                    const size_t sp = p < a.n_edges ? (size_t)a.srcv[p] : 0;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pq + sp * 256 + 128));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pq + sp * 256 + 160));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pq + sp * 256 + 192));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pq + sp * 256 + 224));
                }
            }
            if (pw == 0) IBTL(2, it, 0);                        // rows gathered, waiting for the slot
            if (it > 0) {
                umma::mbar_wait(x_free, (it - 1) & 1);          // the previous tile's MMAs are done with the tile and with R2
                umma::tc_fence_after();
            }
            if (pw == 0) IBTL(2, it, 1);
            regs_to_image(hl);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int64_t p = tile * IB_TE + chalf * 64 + h * 32 + lane;
                const uint32_t dl = p < a.n_edges ? (uint32_t)a.dstv[p] : 0u;
                const uint32_t sl = p < a.n_edges ? (uint32_t)a.srcv[p] : 0u;
                float acc[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = pqn[(size_t)__shfl_sync(0xffffffffu, sl, i) * 256 + 128];
#pragma unroll
                for (int g = 0; g < 32; g += 8) {
                    float pv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) pv[i] = pqn[(size_t)__shfl_sync(0xffffffffu, dl, g + i) * 256];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[g + i] += pv[i];
                }
                umma::tmem_st32(tmem + 128u + lane_base + (uint32_t)(chalf * 64 + h * 32), acc);
            }
            umma::fence_async_smem();
            umma::tc_fence_before();
            if (pw == 0) IBTL(2, it, 2);
            umma::mbar_arrive(x_full);
            if (PASS == 1) {
                // the dz2 tile (aggregation order, whole tiles) replaces the e tile once layer 0 has consumed it
                int64_t row = lane < 16 ? p0 + lane : -1;
                rows_to_regs(a.dz2, row, 1.0f, hl);
                umma::mbar_wait(x_free2, it & 1);
                umma::tc_fence_after();
                regs_to_image(hl);
                umma::fence_async_smem();
                umma::tc_fence_before();
                umma::mbar_arrive(x_full2);
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == IB_MMA_WARP) umma::tmem_dealloc(tmem, 512);
}

// max over nodes and channels of |dagg| / in-degree, as the bits of a non-negative float (integer max: order independent)
__global__ void in_edge_gmax_kernel(const float* __restrict__ dagg, const int32_t* __restrict__ rowptr, int64_t n_nodes, uint32_t* out) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes * 32; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t node = i >> 5;
        const int deg = rowptr[node + 1] - rowptr[node];
        if (deg <= 0) continue;
        const float4 v = reinterpret_cast<const float4*>(dagg)[i];
        const float mx = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) / (float)deg;
        if (mx < 3.0e38f) m = fmaxf(m, mx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// out[i] = sum over the CTAs' partials (fixed order): rows of `len` floats, `stride` floats between CTAs
__global__ void in_edge_partial_reduce_kernel(const float* __restrict__ part, int n_parts, size_t stride, int64_t len, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    float s = 0.f;
    for (int c = 0; c < n_parts; ++c) s += part[(size_t)c * stride + i];
    out[i] = s;
}

static int ib_grid(int64_t n_edges) {
    const int64_t tiles = ceil_div<int64_t>(n_edges, IB_TE);
    return (int)(tiles < sm_count() ? tiles : sm_count());
}

size_t in_edge_bwd_workspace(int64_t n_edges) {
    const int64_t tiles = ceil_div<int64_t>(n_edges > 0 ? n_edges : 1, IB_TE);
    const int64_t sub = ceil_div<int64_t>(n_edges > 0 ? n_edges : 1, IB_FLUSH);
    const size_t grid = (size_t)sm_count();
    return align_up((size_t)tiles * IB_TE * 128 * sizeof(float)) + 2 * align_up((size_t)sub * SEG_H * sizeof(float)) +
           align_up(grid * 4 * 128 * 128 * sizeof(float)) + align_up(grid * 2 * (IB_NVEC_A + IB_NVEC_B) * 128 * sizeof(float)) + 1024;
}

// precision: 2 = plain bf16 (1e-2 contract), anything else = fp16 hi/lo split (1e-5 contract)
int launch_in_edge_bwd(int precision, const float* dagg, const float* e, float e_scale, const int32_t* perm, const float* pq,
                       const int32_t* rowptr, const int32_t* dstv, const int32_t* srcv, int64_t n_nodes, int64_t n_edges,
                       const float* packed, float* dpq, float* dz0, float* dW, float* db, float* dgamma, float* dbeta, int* range_flag,
                       void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(n_edges >= 0 && n_edges < ((int64_t)1 << 31) && n_nodes >= 0 && n_nodes < ((int64_t)1 << 23) * 256,
                "in_edge_bwd: sizes out of range");
    MGB_REQUIRE(((uintptr_t)e % 16) == 0 && ((uintptr_t)pq % 16) == 0 && ((uintptr_t)dagg % 16) == 0, "in_edge_bwd: e / pq / dagg must be 16-byte aligned");
    if (n_nodes > 0) MGB_CUDA(cudaMemsetAsync(dpq, 0, (size_t)n_nodes * 256 * sizeof(float), s));
    MGB_CUDA(cudaMemsetAsync(dW, 0, (size_t)4 * 128 * 128 * sizeof(float), s));
    MGB_CUDA(cudaMemsetAsync(db, 0, (size_t)4 * 128 * sizeof(float), s));
    MGB_CUDA(cudaMemsetAsync(dgamma, 0, 128 * sizeof(float), s));
    MGB_CUDA(cudaMemsetAsync(dbeta, 0, 128 * sizeof(float), s));
    if (n_edges <= 0) return MGB_OK;
    const int64_t tiles = ceil_div<int64_t>(n_edges, IB_TE);
    const int64_t sub = ceil_div<int64_t>(n_edges, IB_FLUSH);
    const int grid = ib_grid(n_edges);
    Workspace ws(ws_ptr, ws_bytes);
    float* dz2 = ws.take<float>((size_t)tiles * IB_TE * 128);
    float* part_head = ws.take<float>((size_t)sub * SEG_H);
    float* part_tail = ws.take<float>((size_t)sub * SEG_H);
    float* wpart = ws.take<float>((size_t)grid * 4 * 128 * 128);
    float* vpart = ws.take<float>((size_t)grid * 2 * (IB_NVEC_A + IB_NVEC_B) * 128);
    uint32_t* gmax = ws.take<uint32_t>(64);
    MGB_WS_CHECK(ws);
    MGB_CUDA(cudaMemsetAsync(wpart, 0, (size_t)grid * 4 * 128 * 128 * sizeof(float), s));
    MGB_CUDA(cudaMemsetAsync(gmax, 0, 64 * sizeof(uint32_t), s));
    {
        const int64_t work = n_nodes * 32;
        const int blocks = (int)(ceil_div<int64_t>(work, 256) < 1184 ? ceil_div<int64_t>(work, 256) : 1184);
        in_edge_gmax_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(dagg, rowptr, n_nodes, gmax);
        MGB_LAUNCH_CHECK();
    }
    InEdgeBwdArgs a{};
    a.e = e; a.e_scale = e_scale; a.perm = perm; a.pq = pq; a.rowptr = rowptr; a.dstv = dstv; a.srcv = srcv; a.n_edges = n_edges;
    a.wimg = packed;
    a.bias = packed + (size_t)5 * 128 * 128;
    a.gamma = a.bias + 5 * 128;
    a.beta = a.gamma + 128;
    a.dagg = dagg; a.gmax_bits = gmax; a.dz2 = dz2; a.dz0 = dz0; a.dpq = dpq; a.part_head = part_head; a.part_tail = part_tail;
    a.range_flag = range_flag;
#ifdef MGB_IB_DEBUG
    { const char* dbg = getenv("MGB_IB_DEBUG"); a.dbg = dbg ? atoi(dbg) : 0; }
#else
    a.dbg = 0;
#endif
    float* vpart_b = vpart + (size_t)grid * 2 * IB_NVEC_A * 128;
    {
        ProfScope prof(PROF_IN_EDGE_BWD, s);
        a.wpart = wpart; a.vpart = vpart;
        if (precision == 2) {
            MGB_CUDA(cudaFuncSetAttribute(in_edge_bwd_tc_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IN_EDGE_BWD_SMEM));
            in_edge_bwd_tc_kernel<0, 1><<<grid, IB_THREADS, IN_EDGE_BWD_SMEM, s>>>(a);
        } else {
            MGB_CUDA(cudaFuncSetAttribute(in_edge_bwd_tc_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IN_EDGE_BWD_SMEM));
            in_edge_bwd_tc_kernel<0, 2><<<grid, IB_THREADS, IN_EDGE_BWD_SMEM, s>>>(a);
        }
        MGB_LAUNCH_CHECK();
        a.wpart = wpart + (size_t)grid * 2 * 128 * 128; a.vpart = vpart_b;
        if (precision == 2) {
            MGB_CUDA(cudaFuncSetAttribute(in_edge_bwd_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IN_EDGE_BWD_SMEM));
            in_edge_bwd_tc_kernel<1, 1><<<grid, IB_THREADS, IN_EDGE_BWD_SMEM, s>>>(a);
        } else {
            MGB_CUDA(cudaFuncSetAttribute(in_edge_bwd_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IN_EDGE_BWD_SMEM));
            in_edge_bwd_tc_kernel<1, 2><<<grid, IB_THREADS, IN_EDGE_BWD_SMEM, s>>>(a);
        }
        MGB_LAUNCH_CHECK();
    }
    MGB_TRY(launch_segment_fixup(rowptr, dstv, n_edges, IB_FLUSH, part_head, part_tail, dpq, 256, 0, s));
    // weight gradients: the accumulators are transposed-free in the partials (partial[n][k] = dW[n][k]); order in dW: W1..W4
    const size_t wstride = (size_t)2 * 128 * 128;
    const int wb = 128 * 128 / 256;
    in_edge_partial_reduce_kernel<<<wb, 256, 0, s>>>(wpart, grid, wstride, 128 * 128, dW + (size_t)3 * 128 * 128);                      // pass A, acc 0: dW4
    MGB_LAUNCH_CHECK();
    in_edge_partial_reduce_kernel<<<wb, 256, 0, s>>>(wpart + 128 * 128, grid, wstride, 128 * 128, dW + (size_t)2 * 128 * 128);          // pass A, acc 1: dW3
    MGB_LAUNCH_CHECK();
    const float* wpb = wpart + (size_t)grid * 2 * 128 * 128;
    in_edge_partial_reduce_kernel<<<wb, 256, 0, s>>>(wpb, grid, wstride, 128 * 128, dW + (size_t)1 * 128 * 128);                        // pass B, acc 0: dW2
    MGB_LAUNCH_CHECK();
    in_edge_partial_reduce_kernel<<<wb, 256, 0, s>>>(wpb + 128 * 128, grid, wstride, 128 * 128, dW);                                    // pass B, acc 1: dW1
    MGB_LAUNCH_CHECK();
    // per-channel gradients: [grid * 2 halves][NV][128]
    float* vdst_a[IB_NVEC_A] = {db + 3 * 128, db + 2 * 128, db + 1 * 128, dgamma, dbeta};
    for (int i = 0; i < IB_NVEC_A; ++i) {
        in_edge_partial_reduce_kernel<<<1, 128, 0, s>>>(vpart + (size_t)i * 128, grid * 2, (size_t)IB_NVEC_A * 128, 128, vdst_a[i]);
        MGB_LAUNCH_CHECK();
    }
    in_edge_partial_reduce_kernel<<<1, 128, 0, s>>>(vpart_b, grid * 2, (size_t)IB_NVEC_B * 128, 128, db);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
