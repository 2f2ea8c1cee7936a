// magnet_b200 — row-wise Linear with the WEIGHTS IN TENSOR MEMORY (TS form of tcgen05.mma), for the two-weight-tile shapes
// of GNN_Layer's node-level stages that linear_tc.cu runs with a single operand stage (its two weight tiles fill shared memory):
//   P | Q = [x, u, pos, var] Wcat^T            128 (+ <= 16 small-K columns) -> 256 outputs     (models/mpnn_2d.py:74-76, factorised)
//   dc    = (d1 * Swish'(y1_pre)) W3           128 -> 256 outputs, residual on the first block  (backward of update_net_1)
//   dx    = dP Wcat[:128] + dQ Wcat[128:]      256 -> 128 outputs, residual                     (backward of the first Linear)
// Same contract as LinTcArgs (internal.cuh) with act = none, no pre-activation copy, all outputs valid.
//   * weights: up to two 128x128 tiles (bf16 hi | lo, 128 TMEM columns each, pack_weight_tmem_bf16) + the small-K columns as one
//     16-wide K-step per output block (16 TMEM columns each), loaded once per CTA; ONE fp32 accumulator (128 columns) that the
//     output blocks pass through in turn;
//   * shared memory is left to the operands: K-major hi | lo images of 128 rows (64 KB), three slots (two + two 32 KB slots for
//     the small-K image when there is one), filled by producer warps a tile ahead;
//   * the epilogue (thread = output channel) first moves its 64 accumulator values to registers and hands the accumulator back,
//     so the MMAs of the next output block / row tile run under its stores: the kernel's only long phase is HBM traffic.
// The small-K columns ride on the tensor core as a ninth K-step instead of 13 FFMAs per output element in the epilogue
// (timeline of the linear_tc.cu version: 26 k cycles of epilogue per 128-row tile of P | Q, 3.5 k of MMA).
#include "internal.cuh"
#include "tc_common.cuh"

namespace mgb {

constexpr int TS_EPI_WARPS = 8, TS_PROD_WARPS = 8, TS_MMA_WARP = 8, TS_PROD_WARP0 = 12;
constexpr int TS_THREADS = (TS_PROD_WARP0 + TS_PROD_WARPS) * 32;      // 640
constexpr uint32_t TS_TMEM_W = 128;        // weight tiles: columns 128 .. 383 (hi | lo, 128 per tile)
constexpr uint32_t TS_TMEM_T = 384;        // small-K weights: 16 columns per output block (hi 8 | lo 8)
constexpr int TS_TAIL_BYTES = 128 * 128;   // one [128 rows][64 columns] bf16 SW128 block (columns 0..15 used)

template <int NSPLIT> constexpr size_t linear_ts_smem() { return 1024 + (size_t)3 * NSPLIT * TILE_BYTES + 256; }

template <int NSPLIT, bool FAST>
__global__ void __launch_bounds__(TS_THREADS, 1) linear_ts_kernel(const LinTcArgs a) {
    constexpr uint32_t SLOT_BYTES = NSPLIT * TILE_BYTES;
    constexpr int NTERM = NSPLIT == 1 ? 1 : 3;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* x_img = base;                                        // [slot][hi|lo]
    unsigned char* t_img = base + (size_t)2 * SLOT_BYTES;               // kt > 0: [slot 0..1][hi|lo] small-K images (in place of slot 2)
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)3 * SLOT_BYTES);
    uint64_t* slot_full = bars;        // [3] producers -> MMA
    uint64_t* slot_free = bars + 3;    // [3] MMA (commit) -> producers
    uint64_t* d_full = bars + 6;       // MMA -> epilogue: an output block is complete
    uint64_t* acc_free = bars + 7;     // epilogue -> MMA: accumulator in registers
    uint64_t* w_full = bars + 8;       // epilogue warps -> MMA: weights in tensor memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.rows, 128);
    const int nt = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);      // row tiles of this CTA (>= 1)
    const int nk = a.nk, nm = a.nm, kt = a.kt;
    const int ring = kt > 0 ? 2 : 3;

    if (tid == 0) {
        for (int q = 0; q < 3; ++q) {
            umma::mbar_init(&slot_full[q], TS_PROD_WARPS * 32);
            umma::mbar_init(&slot_free[q], 1);
        }
        umma::mbar_init(d_full, 1);
        umma::mbar_init(acc_free, TS_EPI_WARPS * 32);
        umma::mbar_init(w_full, TS_EPI_WARPS * 32);
        umma::fence_barrier_init();
    }
    if (warp == TS_MMA_WARP) umma::tmem_alloc(tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < TS_EPI_WARPS) {
        umma::reg_inc<112>();
        // =========================== epilogue: thread = output channel n of the block; warps 0-3 rows 0-63, 4-7 rows 64-127 ====
        const int n = tid & 127, hf = warp >> 2;
        const uint32_t lane_q = (uint32_t)((warp & 3) * 32) << 16;
        // ---- weights -> tensor memory, once: this warp's lane quadrant; hi halves by warps 0-3, lo halves by warps 4-7
        if (hf < NSPLIT) {
#pragma unroll 1
            for (int w = 0; w < nk * nm; ++w) {
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
                    umma::tmem_st32(tmem + lane_q + TS_TMEM_W + (uint32_t)(w * 64 * NSPLIT + hf * 64 + pc * 32), f);
                }
            }
        }
        if (kt > 0 && hf == 0) {
            // the small-K columns of output column m*128 + n: one K-step of 16 (zero-padded), packed pairs, hi 8 | lo 8 columns
#pragma unroll 1
            for (int m = 0; m < nm; ++m) {
                float w[16];
#pragma unroll
                for (int t = 0; t < 16; ++t) w[t] = t < kt ? a.wtail[(int64_t)(m * 128 + n) * a.wt_sn + (int64_t)t * a.wt_st] : 0.f;
                float f[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t hi, lo;
                    split2_bf16(w[2 * j], w[2 * j + 1], hi, lo);
                    f[j] = __uint_as_float(hi);
                    f[8 + j] = __uint_as_float(NSPLIT == 2 ? lo : 0u);
                }
                umma::tmem_st16(tmem + lane_q + TS_TMEM_T + (uint32_t)(m * 16), f);
            }
        }
        umma::tc_fence_before();
        umma::mbar_arrive(w_full);
        float bias_m[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) bias_m[m] = (a.bias && m < nm) ? a.bias[m * 128 + n] : 0.f;
        const int ldy = a.ldy, ldr = a.ldr;
        uint32_t j = 0;               // output blocks completed so far
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 128;
            const int nr = (int)((a.rows - r0) < 128 ? (a.rows - r0) : 128);
            const int rlim = nr - hf * 64;                   // valid rows among the 64 of this half
            const int64_t rb = r0 + hf * 64;
            const uint32_t tacc = tmem + lane_q + (uint32_t)(hf * 64);
#pragma unroll 1
            for (int m = 0; m < nm; ++m, ++j) {
                const int col = m * 128 + n;
                const float bias = m ? bias_m[1] : bias_m[0];
                const bool has_res = a.residual && (a.res_blocks == 0 || ((a.res_blocks >> m) & 1));
                const float* rp = has_res ? a.residual + rb * ldr + col : nullptr;
                float* yo = a.y + rb * ldy + col;
                float res[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {       // in flight while the MMAs run
                    res[i] = 0.f;
                    if (has_res) {
                        if (i < rlim) res[i] = *rp;
                        rp += ldr;
                    }
                }
                umma::mbar_wait(d_full, j & 1);
                umma::tc_fence_after();
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
                    for (int i = 0; i < 16; ++i) v[c + i] += bias + res[i];
                    if (c + 16 < 64) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            res[i] = 0.f;
                            if (has_res) {
                                if (c + 16 + i < rlim) res[i] = *rp;
                                rp += ldr;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (i < lim) *yo = v[c + i];
                        yo += ldy;
                    }
                }
            }
        }
    } else if (warp < TS_PROD_WARP0 && warp != TS_MMA_WARP) {
        umma::reg_dec<40>();          // padding warps of the MMA warpgroup
    } else if (warp == TS_MMA_WARP) {
        umma::reg_dec<40>();
        // =========================== MMA issue =======================================================
        const uint32_t id_k = umma::idesc_bf16(128, 128, 0, 0);
        const uint64_t xk_d = umma::desc_sw128(umma::smem_u32(x_img), 16, 1024);
        const uint64_t tk_d = umma::desc_sw128(umma::smem_u32(t_img), 16, 1024);
        constexpr uint32_t TB = TILE_BYTES >> 4, SB = SLOT_BYTES >> 4, TTB = TS_TAIL_BYTES >> 4;
        const uint32_t d = tmem;
        umma::mbar_wait(w_full, 0);
        umma::tc_fence_after();
        uint32_t j = 0;
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
#pragma unroll 1
            for (int m = 0; m < nm; ++m, ++j) {
#pragma unroll 1
                for (int kc = 0; kc < nk; ++kc) {
                    const int q = it * nk + kc, slot = q % ring;
                    if (m == 0) umma::mbar_wait(&slot_full[slot], (q / ring) & 1);
                    if (kc == 0) umma::mbar_wait(acc_free, (j & 1) ^ 1);      // accumulator taken over by the previous block's epilogue
                    umma::tc_fence_after();
                    if (umma::elect_one()) {
                        const uint32_t w_t = tmem + TS_TMEM_W + (uint32_t)(a.tile_of[m][kc] * 64 * NSPLIT);
                        const uint64_t xd = xk_d + (uint64_t)((uint32_t)slot * SB);
#pragma unroll
                        for (int term = 0; term < NTERM; ++term) {            // hi*hi, hi*lo, lo*hi
                            const uint32_t wa = w_t + (term == 2 ? 64u : 0u);
                            const uint64_t xb = xd + (term == 1 ? TB : 0);
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                umma::mma_bf16_ts(d, wa + (uint32_t)(k * 8), xb + (uint64_t)((k >> 2) * (128 * 128 >> 4) + (k & 3) * 2), id_k,
                                                  (kc | term | k) ? 1u : 0u);
                        }
                        if (kt > 0) {      // ninth K-step: the small-K columns
                            const uint32_t tw = tmem + TS_TMEM_T + (uint32_t)(m * 16);
                            const uint64_t td = tk_d + (uint64_t)((uint32_t)slot * NSPLIT * TTB);
#pragma unroll
                            for (int term = 0; term < NTERM; ++term)
                                umma::mma_bf16_ts(d, tw + (term == 2 ? 8u : 0u), td + (term == 1 ? TTB : 0), id_k, 1u);
                        }
                        if (m == nm - 1) umma::mma_commit(&slot_free[slot]);
                        if (kc == nk - 1) umma::mma_commit(d_full);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        umma::reg_inc<104>();
        // =========================== producers: 16 rows per warp, one K-tile of a row tile after the other ===========
        const int pw = warp - TS_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
#pragma unroll 1
        for (int q = 0; q < nt * nk; ++q) {
            const int it = q / nk, kc = q - it * nk, slot = q % ring;
            const float* src = a.src[kc];
            const int ld = a.ld[kc];
            const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 128 + pw * 16;
            const int64_t left = a.rows - r0;
            const int nv = left >= 16 ? 16 : (left > 0 ? (int)left : 0);       // valid rows of this warp
            const int64_t rfirst = nv > 0 ? r0 : a.rows - 1;
            {   // L2 prefetch of this warp's rows of the CTA's next tile (same source)
                const int64_t rn = r0 + (int64_t)gridDim.x * 128;
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int line = jj * 32 + lane;
                    const int64_t row = rn + (line >> 2);
                    if (row < a.rows) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + row * ld + (line & 3) * 32));
                        if (kc == 0 && a.pre) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pre + row * a.ldpre + (line & 3) * 32));
                    }
                }
            }
            float4 x[16];
            {
                const float* p = src + rfirst * ld + lane * 4;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    x[r] = *reinterpret_cast<const float4*>(p);
                    if (r + 1 < nv) p += ld;
                }
            }
            if (kc == 0 && a.pre) {
                // pre-activation rows in two halves of eight, the first requested together with the x rows
                const float* p = a.pre + rfirst * a.ldpre + lane * 4;
                const int ldp = a.ldpre;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 qv[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        qv[r] = *reinterpret_cast<const float4*>(p);
                        if (h * 8 + r + 1 < nv) p += ldp;
                    }
                    if (a.pre_act == ACT_SWISH) {
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            x[h * 8 + r].x *= swish_grad_tc<FAST>(qv[r].x);
                            x[h * 8 + r].y *= swish_grad_tc<FAST>(qv[r].y);
                            x[h * 8 + r].z *= swish_grad_tc<FAST>(qv[r].z);
                            x[h * 8 + r].w *= swish_grad_tc<FAST>(qv[r].w);
                        }
                    } else if (a.pre_act == ACT_RELU) {
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            x[h * 8 + r].x = qv[r].x > 0.f ? x[h * 8 + r].x : 0.f;
                            x[h * 8 + r].y = qv[r].y > 0.f ? x[h * 8 + r].y : 0.f;
                            x[h * 8 + r].z = qv[r].z > 0.f ? x[h * 8 + r].z : 0.f;
                            x[h * 8 + r].w = qv[r].w > 0.f ? x[h * 8 + r].w : 0.f;
                        }
                    }
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
            // small-K image of the row tile (kc == 0 only when kt > 0): lane -> row pw*16 + lane/2, columns (lane & 1)*8 .. +7
            uint4 thi = make_uint4(0u, 0u, 0u, 0u), tlo = make_uint4(0u, 0u, 0u, 0u);
            if (kt > 0) {
                const int rr = lane >> 1;
                float tv[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int t = (lane & 1) * 8 + jj;
                    tv[jj] = 0.f;
                    if (t < kt && rr < nv) {
                        int tt = t, seg = 0;
                        while (tt >= a.tk[seg]) { tt -= a.tk[seg]; ++seg; }
                        tv[jj] = a.tsrc[seg][(r0 + rr) * a.tld[seg] + tt];
                    }
                }
                split2_bf16(tv[0], tv[1], thi.x, tlo.x);
                split2_bf16(tv[2], tv[3], thi.y, tlo.y);
                split2_bf16(tv[4], tv[5], thi.z, tlo.z);
                split2_bf16(tv[6], tv[7], thi.w, tlo.w);
            }
            umma::mbar_wait(&slot_free[slot], ((q / ring) & 1) ^ 1);
            unsigned char* img = x_img + (size_t)slot * SLOT_BYTES;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                *reinterpret_cast<uint2*>(img + off) = make_uint2(hl[r].x, hl[r].y);
                if (NSPLIT == 2) *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = make_uint2(hl[r].z, hl[r].w);
            }
            if (kt > 0) {
                const int r = pw * 16 + (lane >> 1);
                unsigned char* ti = t_img + (size_t)slot * NSPLIT * TS_TAIL_BYTES;
                const uint32_t off = (uint32_t)r * 128u + (uint32_t)(((lane & 1) ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(ti + off) = thi;
                if (NSPLIT == 2) *reinterpret_cast<uint4*>(ti + TS_TAIL_BYTES + off) = tlo;
            }
            umma::fence_async_smem();
            umma::mbar_arrive(&slot_full[slot]);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == TS_MMA_WARP) umma::tmem_dealloc(tmem, 512);
}

// true when launch_linear_ts covers the call (the caller then passes TMEM-format weight images in a.wimg)
bool linear_ts_covers(const LinTcArgs& a) {
    return a.nk >= 1 && a.nk <= 2 && a.nm >= 1 && a.nm <= 2 && a.nk * a.nm <= 2 && a.kt >= 0 && a.kt <= 16 && (a.kt == 0 || a.nk == 1) &&
           a.act == ACT_NONE && a.y_pre == nullptr && a.self_act == ACT_NONE && a.n_out == 0 && a.a_trans == 0;
}

int launch_linear_ts(int precision, const LinTcArgs& a, cudaStream_t s) {
    MGB_REQUIRE(linear_ts_covers(a), "linear_ts: shape / epilogue not covered");
    MGB_REQUIRE(a.rows < ((int64_t)1 << 31), "linear_ts: row count out of range");
    for (int kc = 0; kc < a.nk; ++kc)
        MGB_REQUIRE(((uintptr_t)a.src[kc] % 16) == 0 && a.ld[kc] % 4 == 0, "linear_ts: source rows must be 16-byte aligned");
    MGB_REQUIRE(a.pre == nullptr || (((uintptr_t)a.pre % 16) == 0 && a.ldpre % 4 == 0), "linear_ts: pre-activation rows must be 16-byte aligned");
    MGB_REQUIRE(((uintptr_t)a.wimg % 16) == 0, "linear_ts: packed weights must be 16-byte aligned");
    if (a.rows <= 0) return MGB_OK;
    const int64_t tiles = ceil_div<int64_t>(a.rows, 128);
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    ProfScope prof(PROF_NODE_GEMM, s);
    if (precision == 2) {
        MGB_CUDA(cudaFuncSetAttribute(linear_ts_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)linear_ts_smem<1>()));
        linear_ts_kernel<1, true><<<grid, TS_THREADS, linear_ts_smem<1>(), s>>>(a);
    } else {
        MGB_CUDA(cudaFuncSetAttribute(linear_ts_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)linear_ts_smem<2>()));
        linear_ts_kernel<2, false><<<grid, TS_THREADS, linear_ts_smem<2>(), s>>>(a);
    }
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
