// magnet_b200 — row-wise Linear layers on the tensor cores (tcgen05 + TMEM): the node-level stages of GNN_Layer
// (first-Linear factorisation P|Q, update_net_1/2 — models/mpnn_2d.py:51-62,85-90), their data gradients, and the
// weight gradients; also the 128-wide Linears of MLP (models/backbones/mlp.py:9-28).
//
// Same transposed formulation as the edge kernels:  D^T[n][row] = sum_k W[n][k] X[row][k]
//   A = weight tile(s) (M = 128 output channels per block, resident in shared memory as pre-packed swizzled images),
//   B = a tile of 128 rows of X (N), built by producer warps from fp32 rows (optionally transformed on the fly:
//       x * act'(pre) for data gradients, act(x) for recomputed activations), converted to bf16 hi/lo,
//   D^T in TMEM; an epilogue thread owns one output channel: + bias + small-K tail (FFMA) -> activation -> + residual,
//   one coalesced 128-byte store per warp and row.
// K = 128*nk (+ a tail of <= 16 columns evaluated in the epilogue), outputs = 128*nm (nm <= 2), nm*nk <= 2 weight tiles.
// Data gradients use the SAME weight images through MN-major descriptors (no transposed copy).
#include "internal.cuh"
#include "tc_common.cuh"

namespace mgb {

constexpr int LT_EPI_WARPS = 4, LT_PROD_WARPS = 8;
constexpr int LT_MMA_WARP = LT_EPI_WARPS, LT_PROD_WARP0 = LT_EPI_WARPS + 1;
constexpr int LT_THREADS = (LT_EPI_WARPS + 1 + LT_PROD_WARPS) * 32;   // 416
constexpr int LT_TAIL = 16;

// W [rows_total][ld] fp32, tile = W[r0:r0+128, c0:c0+128] (zero outside) -> swizzled bf16 images hi | lo
__global__ void pack_weight_tile_kernel(const float* __restrict__ W, int ld, int n_rows, int n_cols, int r0, int c0,
                                        unsigned char* __restrict__ img) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * 128) return;
    const int r = idx >> 7, c = idx & 127;
    const float v = (r0 + r < n_rows && c0 + c < n_cols) ? W[(int64_t)(r0 + r) * ld + c0 + c] : 0.f;
    __nv_bfloat16 hi, lo;
    umma::split_bf16(v, hi, lo);
    const uint32_t off = umma::tile_off(128, r, c);
    *reinterpret_cast<__nv_bfloat16*>(img + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(img + TILE_BYTES + off) = lo;
}

int pack_weight_tile(const float* W, int ld, int n_rows, int n_cols, int r0, int c0, void* img, cudaStream_t s) {
    pack_weight_tile_kernel<<<64, 256, 0, s>>>(W, ld, n_rows, n_cols, r0, c0, (unsigned char*)img);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// 2 weight tiles (always hi|lo) + B stages (2 x bf16 or 1 x hi|lo) + 2 tail tiles
constexpr size_t LINEAR_TC_SMEM = 1024 + (size_t)4 * TILE_BYTES + (size_t)2 * TILE_BYTES + 2 * 128 * LT_TAIL * sizeof(float) + 256;

template <int NSPLIT, bool FAST>
__global__ void __launch_bounds__(LT_THREADS, 1) linear_tc_kernel(const LinTcArgs a) {
    constexpr int BSTAGES = NSPLIT == 1 ? 2 : 1;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* w_img = base;                                        // [2 tiles][hi|lo]
    unsigned char* b_img = w_img + (size_t)4 * TILE_BYTES;               // [BSTAGES][NSPLIT]
    float* tails = reinterpret_cast<float*>(b_img + (size_t)BSTAGES * NSPLIT * TILE_BYTES);   // [2][128][LT_TAIL]
    uint64_t* bars = reinterpret_cast<uint64_t*>(tails + 2 * 128 * LT_TAIL);
    uint64_t* full = bars;            // [2]
    uint64_t* empty = bars + 2;       // [2]
    uint64_t* tfull = bars + 4;       // [2] accumulators ready
    uint64_t* tempty = bars + 6;      // [2] accumulators (and tail tile) drained
    uint64_t* tailfull = bars + 8;    // [2] tail tile written
    uint64_t* wbar = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.rows, 128);
    const int n_wtiles = a.nm * a.nk;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&full[s], LT_PROD_WARPS * 32);
            umma::mbar_init(&empty[s], 1);
            umma::mbar_init(&tfull[s], 1);
            umma::mbar_init(&tempty[s], LT_EPI_WARPS * 32);
            umma::mbar_init(&tailfull[s], LT_PROD_WARPS * 32);
        }
        umma::mbar_init(wbar, 1);
        umma::fence_barrier_init();
    }
    if (warp == LT_MMA_WARP) umma::tmem_alloc(tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < LT_EPI_WARPS) {
        // =========================== epilogue: thread = output channel within the block =================
        const int n = tid;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int64_t r0 = tile * 128;
            const int nr = (int)((a.rows - r0) < 128 ? (a.rows - r0) : 128);
            const float* tl = tails + acc * 128 * LT_TAIL;
            if (a.kt > 0) umma::mbar_wait(&tailfull[acc], aph);
            umma::mbar_wait(&tfull[acc], aph);
            umma::tc_fence_after();
            for (int m = 0; m < a.nm; ++m) {
                const int col = m * 128 + n;
                const float bias = a.bias ? a.bias[col] : 0.f;
                float wt[LT_TAIL];
#pragma unroll
                for (int t = 0; t < LT_TAIL; ++t) wt[t] = t < a.kt ? a.wtail[(int64_t)col * a.wt_sn + (int64_t)t * a.wt_st] : 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    float v[32];
                    umma::tmem_ld32(tmem + (uint32_t)((acc * 2 + m) * 128) + ((uint32_t)(warp * 32) << 16) + c0, v);
                    if (m == a.nm - 1 && c0 + 32 >= 128 && a.kt == 0) {
                        umma::tc_fence_before();
                        umma::mbar_arrive(&tempty[acc]);
                    }
                    if (c0 >= nr) continue;
                    if (a.kt > 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float4* trow = reinterpret_cast<const float4*>(tl + (c0 + i) * LT_TAIL);
                            float s = v[i];
#pragma unroll
                            for (int t4 = 0; t4 < LT_TAIL / 4; ++t4) {
                                if (t4 * 4 < a.kt) {
                                    const float4 x = trow[t4];
                                    s = fmaf(x.x, wt[t4 * 4 + 0], s);
                                    s = fmaf(x.y, wt[t4 * 4 + 1], s);
                                    s = fmaf(x.z, wt[t4 * 4 + 2], s);
                                    s = fmaf(x.w, wt[t4 * 4 + 3], s);
                                }
                            }
                            v[i] = s;
                        }
                    }
                    const int lim = nr - c0 < 32 ? nr - c0 : 32;
                    if (a.y_pre) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < lim) a.y_pre[(r0 + c0 + i) * a.ldyp + col] = v[i] + bias;
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = act_tc<FAST>(a.act, v[i] + bias);
                    if (a.residual) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < lim) v[i] += a.residual[(r0 + c0 + i) * a.ldr + col];
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < lim) a.y[(r0 + c0 + i) * a.ldy + col] = v[i];
                }
            }
            if (a.kt > 0) {          // the tail tile is read until the end: release accumulators and tail together
                umma::tc_fence_before();
                umma::mbar_arrive(&tempty[acc]);
            }
        }
    } else if (warp == LT_MMA_WARP) {
        // =========================== MMA issue ====================================================
        if (lane == 0) load_w2_image(w_img, (const unsigned char*)a.wimg, (uint32_t)(n_wtiles * 2 * TILE_BYTES), wbar);   // global images always hold hi|lo
        umma::mbar_wait(wbar, 0);
        const uint32_t idesc = umma::idesc_bf16(128, 128, a.a_trans, 0);
        const uint32_t w_s = umma::smem_u32(w_img);
        int it = 0, sc = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            umma::mbar_wait(&tempty[acc], aph ^ 1);
            for (int kc = 0; kc < a.nk; ++kc, ++sc) {
                const int s = sc % BSTAGES;
                umma::mbar_wait(&full[s], (sc / BSTAGES) & 1);
                umma::tc_fence_after();
                if (lane == 0) {
                    const uint32_t b_s = umma::smem_u32(b_img + (size_t)s * NSPLIT * TILE_BYTES);
                    for (int m = 0; m < a.nm; ++m) {
                        const uint32_t d = tmem + (uint32_t)((acc * 2 + m) * 128);
                        const uint32_t wt_s = w_s + (uint32_t)a.tile_of[m][kc] * 2 * TILE_BYTES;
                        uint32_t accum = kc > 0 ? 1u : 0u;
#pragma unroll
                        for (int term = 0; term < (NSPLIT == 1 ? 1 : 3); ++term) {
                            const int wa = term == 2 ? 1 : 0, hb = term == 1 ? 1 : 0;
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const uint64_t da = a.a_trans ? umma::desc_mnmajor(wt_s + wa * TILE_BYTES, k)
                                                              : umma::desc_kmajor(wt_s + wa * TILE_BYTES, k);
                                umma::mma_bf16(d, da, umma::desc_kmajor(b_s + hb * TILE_BYTES, k), idesc, accum);
                                accum = 1;
                            }
                        }
                    }
                    umma::mma_commit(&empty[s]);
                    if (kc == a.nk - 1) umma::mma_commit(&tfull[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================== producers: 16 rows per warp ===================================
        const int pw = warp - LT_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        int it = 0, sc = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int64_t r0 = tile * 128 + pw * 16;
            if (a.kt > 0) {
                umma::mbar_wait(&tempty[acc], aph ^ 1);
                float* tl = tails + acc * 128 * LT_TAIL + pw * 16 * LT_TAIL;
                // lane -> (row = lane/2 + 16*half..., ) : 16 rows x 16 tail slots = 256 values, 8 per lane
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int e = j * 32 + lane, r = e >> 4, t = e & 15;
                    float v = 0.f;
                    if (r0 + r < a.rows && t < a.kt) {
                        int tt = t, seg = 0;
                        while (tt >= a.tk[seg]) { tt -= a.tk[seg]; ++seg; }
                        v = a.tsrc[seg][(r0 + r) * a.tld[seg] + tt];
                    }
                    tl[r * LT_TAIL + t] = v;
                }
                umma::mbar_arrive(&tailfull[acc]);
            }
            for (int kc = 0; kc < a.nk; ++kc, ++sc) {
                const int s = sc % BSTAGES;
                const float* src = a.src[kc];
                const int ld = a.ld[kc];
                float4 x[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int64_t row = r0 + r < a.rows ? r0 + r : a.rows - 1;
                    x[r] = *reinterpret_cast<const float4*>(src + row * ld + lane * 4);
                }
                if (kc == 0 && a.pre) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const int64_t row = r0 + r < a.rows ? r0 + r : a.rows - 1;
                        const float4 p = *reinterpret_cast<const float4*>(a.pre + row * a.ldpre + lane * 4);
                        x[r].x *= act_grad_tc<FAST>(a.pre_act, p.x);
                        x[r].y *= act_grad_tc<FAST>(a.pre_act, p.y);
                        x[r].z *= act_grad_tc<FAST>(a.pre_act, p.z);
                        x[r].w *= act_grad_tc<FAST>(a.pre_act, p.w);
                    }
                }
                if (kc == 0 && a.self_act) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        x[r].x = act_tc<FAST>(a.self_act, x[r].x);
                        x[r].y = act_tc<FAST>(a.self_act, x[r].y);
                        x[r].z = act_tc<FAST>(a.self_act, x[r].z);
                        x[r].w = act_tc<FAST>(a.self_act, x[r].w);
                    }
                }
                umma::mbar_wait(&empty[s], ((sc / BSTAGES) & 1) ^ 1);
                unsigned char* img = b_img + (size_t)s * NSPLIT * TILE_BYTES;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    float4 h = x[r];
                    if (r0 + r >= a.rows) h = make_float4(0.f, 0.f, 0.f, 0.f);
                    const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                    if (NSPLIT == 1) {
                        *reinterpret_cast<uint2*>(img + off) = make_uint2(umma::pack_bf16(h.x, h.y), umma::pack_bf16(h.z, h.w));
                    } else {
                        uint2 hi, lo;
                        split2_bf16(h.x, h.y, hi.x, lo.x);
                        split2_bf16(h.z, h.w, hi.y, lo.y);
                        *reinterpret_cast<uint2*>(img + off) = hi;
                        *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = lo;
                    }
                }
                umma::fence_async_smem();
                umma::mbar_arrive(&full[s]);
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == LT_MMA_WARP) umma::tmem_dealloc(tmem, 512);
}

int launch_linear_tc(int precision, const LinTcArgs& a, cudaStream_t s) {
    MGB_REQUIRE(a.nk >= 1 && a.nk <= 2 && a.nm >= 1 && a.nm <= 2 && a.nm * a.nk <= 2, "linear_tc: at most two weight tiles");
    MGB_REQUIRE(a.kt >= 0 && a.kt <= LT_TAIL, "linear_tc: tail width must be <= %d", LT_TAIL);
    MGB_REQUIRE(a.rows < ((int64_t)1 << 31), "linear_tc: row count out of range");
    if (a.rows <= 0) return MGB_OK;
    const int64_t tiles = ceil_div<int64_t>(a.rows, 128);
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    ProfScope prof(PROF_NODE_GEMM, s);
    if (precision == 2) {
        MGB_CUDA(cudaFuncSetAttribute(linear_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LINEAR_TC_SMEM));
        linear_tc_kernel<1, true><<<grid, LT_THREADS, LINEAR_TC_SMEM, s>>>(a);
    } else {
        MGB_CUDA(cudaFuncSetAttribute(linear_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LINEAR_TC_SMEM));
        linear_tc_kernel<2, false><<<grid, LT_THREADS, LINEAR_TC_SMEM, s>>>(a);
    }
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ==================================================================================================
// Weight gradient on the tensor cores:  dW[n][k] (+)= sum_rows Y'[row][n] X[row][k]   (one 128x128 tile per launch)
//   Y' = dy * act'(y_pre) (optional), X = a 128-column source (optionally act(x)) or the packed small-K tail columns.
// Both operands are row tiles [row][col] used as MN-major operands (K = rows); the accumulator lives in TMEM across
// the CTA's row tiles, two buffers alternate every 4 tiles and are drained with round-to-nearest adds (see gnn_edge_tc.cu).
// ==================================================================================================
constexpr size_t WGRAD_TC_SMEM = 1024 + (size_t)4 * TILE_BYTES + 256;

template <int NSPLIT, bool FAST>
__global__ void __launch_bounds__(LT_THREADS, 1) wgrad_tc_kernel(const WgradTcArgs a) {
    constexpr int GROUP = FAST ? (1 << 30) : 4;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* y_img = base;                                  // [NSPLIT]  Y'[row][n]
    unsigned char* x_img = base + (size_t)2 * TILE_BYTES;         // [NSPLIT]  X[row][k]
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)4 * TILE_BYTES);
    uint64_t* full = bars;          // producers -> MMA
    uint64_t* empty = bars + 1;     // MMA -> producers
    uint64_t* d_full = bars + 2;    // [2]
    uint64_t* d_empty = bars + 4;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.rows, 128);
    if (tid == 0) {
        umma::mbar_init(full, LT_PROD_WARPS * 32);
        umma::mbar_init(empty, 1);
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&d_full[s], 1);
            umma::mbar_init(&d_empty[s], LT_EPI_WARPS * 32);
        }
        umma::fence_barrier_init();
    }
    if (warp == LT_MMA_WARP) umma::tmem_alloc(tmem_slot, 256);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < LT_EPI_WARPS) {
        const int n = tid;
        float* out = a.partial + ((int64_t)blockIdx.x * 128 + n) * 128;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const bool last = tile + gridDim.x >= n_tiles;
            if ((it % GROUP) != GROUP - 1 && !last) continue;
            const int grp = it / GROUP, buf = grp & 1;
            umma::mbar_wait(&d_full[buf], (grp >> 1) & 1);
            umma::tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float v[32];
                umma::tmem_ld32(tmem + (uint32_t)(buf * 128) + ((uint32_t)(warp * 32) << 16) + c0, v);
                if (c0 + 32 >= 128) {
                    umma::tc_fence_before();
                    umma::mbar_arrive(&d_empty[buf]);
                }
                float4* o = reinterpret_cast<float4*>(out + c0);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    float4 w = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
                    if (grp > 0) {
                        const float4 old = o[q4];
                        w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
                    }
                    o[q4] = w;
                }
            }
        }
    } else if (warp == LT_MMA_WARP) {
        const uint32_t idesc = umma::idesc_bf16(128, 128, 1, 1);
        const uint32_t y_s = umma::smem_u32(y_img), x_s = umma::smem_u32(x_img);
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int grp = it / GROUP, buf = grp & 1;
            const bool first_in_group = (it % GROUP) == 0;
            const bool last = tile + gridDim.x >= n_tiles;
            if (first_in_group) umma::mbar_wait(&d_empty[buf], ((grp >> 1) & 1) ^ 1);
            umma::mbar_wait(full, it & 1);
            umma::tc_fence_after();
            if (lane == 0) {
                uint32_t accum = first_in_group ? 0u : 1u;
#pragma unroll
                for (int term = 0; term < (NSPLIT == 1 ? 1 : 3); ++term) {
                    const int ya = term == 2 ? 1 : 0, xb = term == 1 ? 1 : 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        umma::mma_bf16(tmem + (uint32_t)(buf * 128), umma::desc_mnmajor(y_s + ya * TILE_BYTES, k),
                                       umma::desc_mnmajor(x_s + xb * TILE_BYTES, k), idesc, accum);
                        accum = 1;
                    }
                }
                umma::mma_commit(empty);
                if ((it % GROUP) == GROUP - 1 || last) umma::mma_commit(&d_full[buf]);
            }
            __syncwarp();
        }
    } else {
        const int pw = warp - LT_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int64_t r0 = tile * 128 + pw * 16;
            float4 yv[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int64_t row = r0 + r < a.rows ? r0 + r : a.rows - 1;
                yv[r] = *reinterpret_cast<const float4*>(a.dy + row * a.lddy + lane * 4);
            }
            if (a.y_pre) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int64_t row = r0 + r < a.rows ? r0 + r : a.rows - 1;
                    const float4 p = *reinterpret_cast<const float4*>(a.y_pre + row * a.ldyp + lane * 4);
                    yv[r].x *= act_grad_tc<FAST>(a.y_act, p.x);
                    yv[r].y *= act_grad_tc<FAST>(a.y_act, p.y);
                    yv[r].z *= act_grad_tc<FAST>(a.y_act, p.z);
                    yv[r].w *= act_grad_tc<FAST>(a.y_act, p.w);
                }
            }
            umma::mbar_wait(empty, (it & 1) ^ 1);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                float4 h = yv[r];
                if (r0 + r >= a.rows) h = make_float4(0.f, 0.f, 0.f, 0.f);
                const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                if (NSPLIT == 1) {
                    *reinterpret_cast<uint2*>(y_img + off) = make_uint2(umma::pack_bf16(h.x, h.y), umma::pack_bf16(h.z, h.w));
                } else {
                    uint2 hi, lo;
                    split2_bf16(h.x, h.y, hi.x, lo.x);
                    split2_bf16(h.z, h.w, hi.y, lo.y);
                    *reinterpret_cast<uint2*>(y_img + off) = hi;
                    *reinterpret_cast<uint2*>(y_img + TILE_BYTES + off) = lo;
                }
            }
            // X tile: a 128-column source, or the small-K tail columns packed into columns [0, kt)
#pragma unroll 4
            for (int r = 0; r < 16; ++r) {
                float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + r < a.rows) {
                    if (a.x) {
                        h = *reinterpret_cast<const float4*>(a.x + (r0 + r) * a.ldx + lane * 4);
                        if (a.x_act) {
                            h.x = act_tc<FAST>(a.x_act, h.x); h.y = act_tc<FAST>(a.x_act, h.y);
                            h.z = act_tc<FAST>(a.x_act, h.z); h.w = act_tc<FAST>(a.x_act, h.w);
                        }
                    } else {
                        float t4[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int t = lane * 4 + u;
                            float v = 0.f;
                            if (t < a.kt) {
                                int tt = t, seg = 0;
                                while (tt >= a.tk[seg]) { tt -= a.tk[seg]; ++seg; }
                                v = a.tsrc[seg][(r0 + r) * a.tld[seg] + tt];
                            }
                            t4[u] = v;
                        }
                        h = make_float4(t4[0], t4[1], t4[2], t4[3]);
                    }
                }
                const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                if (NSPLIT == 1) {
                    *reinterpret_cast<uint2*>(x_img + off) = make_uint2(umma::pack_bf16(h.x, h.y), umma::pack_bf16(h.z, h.w));
                } else {
                    uint2 hi, lo;
                    split2_bf16(h.x, h.y, hi.x, lo.x);
                    split2_bf16(h.z, h.w, hi.y, lo.y);
                    *reinterpret_cast<uint2*>(x_img + off) = hi;
                    *reinterpret_cast<uint2*>(x_img + TILE_BYTES + off) = lo;
                }
            }
            umma::fence_async_smem();
            umma::mbar_arrive(full);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == LT_MMA_WARP) umma::tmem_dealloc(tmem, 256);
}

// dw[(n0+n)*lddw + k0 + k] (+)= sum over CTA partials, for n < n_valid, k < k_valid
__global__ void __launch_bounds__(256)
wgrad_tc_reduce_kernel(const float* __restrict__ partial, int n_parts, int n_valid, int k_valid, float* __restrict__ dw, int lddw,
                       int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * 128) return;
    const int n = idx >> 7, k = idx & 127;
    if (n >= n_valid || k >= k_valid) return;
    float s = 0.f;
    for (int p = 0; p < n_parts; ++p) s += partial[(int64_t)p * 128 * 128 + idx];
    float* o = dw + (int64_t)n * lddw + k;
    *o = accumulate ? *o + s : s;
}

static int wgrad_tc_grid(int64_t rows) {
    const int64_t tiles = ceil_div<int64_t>(rows > 0 ? rows : 1, 128);
    return (int)(tiles < sm_count() ? tiles : sm_count());
}
size_t wgrad_tc_workspace(int64_t rows) { return align_up((size_t)wgrad_tc_grid(rows) * 128 * 128 * sizeof(float)) + 256; }

int launch_wgrad_tc(int precision, WgradTcArgs a, float* dw, int lddw, int n_valid, int k_valid, int accumulate, void* ws_ptr,
                    size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(a.rows < ((int64_t)1 << 31), "wgrad_tc: row count out of range");
    if (a.rows <= 0) return MGB_OK;
    const int grid = wgrad_tc_grid(a.rows);
    Workspace ws(ws_ptr, ws_bytes);
    a.partial = ws.take<float>((size_t)grid * 128 * 128);
    MGB_WS_CHECK(ws);
    {
        ProfScope prof(PROF_WGRAD, s);
        if (precision == 2) {
            MGB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WGRAD_TC_SMEM));
            wgrad_tc_kernel<1, true><<<grid, LT_THREADS, WGRAD_TC_SMEM, s>>>(a);
        } else {
            MGB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WGRAD_TC_SMEM));
            wgrad_tc_kernel<2, false><<<grid, LT_THREADS, WGRAD_TC_SMEM, s>>>(a);
        }
    }
    MGB_LAUNCH_CHECK();
    wgrad_tc_reduce_kernel<<<64, 256, 0, s>>>(a.partial, grid, n_valid, k_valid, dw, lddw, accumulate);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// out[row][t] = sum_n Y'[row][n] * Wt(n, t),  t < kt <= 16, Y' = dy * act'(pre);  ny = 128 or 256 columns.
// One warp per row (8 rows per block, blocks stride over the rows); the tail weights sit in shared memory as [n][16].
__global__ void __launch_bounds__(256)
tail_dgrad_kernel(const float* __restrict__ dy, int lddy, int ny, const float* __restrict__ pre, int ldpre, int act,
                  const float* __restrict__ wtail, int wt_sn, int wt_st, int kt, int64_t rows, float* __restrict__ out, int ldo) {
    __shared__ __align__(16) float wts[256 * LT_TAIL];
    for (int i = threadIdx.x; i < ny * LT_TAIL; i += blockDim.x) {
        const int n = i / LT_TAIL, t = i % LT_TAIL;
        wts[i] = t < kt ? wtail[(int64_t)n * wt_sn + (int64_t)t * wt_st] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int64_t row = (int64_t)blockIdx.x * 8 + wib; row < rows; row += (int64_t)gridDim.x * 8) {
        float acc[LT_TAIL];
#pragma unroll
        for (int t = 0; t < LT_TAIL; ++t) acc[t] = 0.f;
        for (int c0 = 0; c0 < ny; c0 += 128) {
            const int n0 = c0 + lane * 4;
            const float4 g = *reinterpret_cast<const float4*>(dy + row * lddy + n0);
            float y[4] = {g.x, g.y, g.z, g.w};
            if (pre) {
                const float4 p = *reinterpret_cast<const float4*>(pre + row * ldpre + n0);
                y[0] *= act_grad(act, p.x); y[1] *= act_grad(act, p.y); y[2] *= act_grad(act, p.z); y[3] *= act_grad(act, p.w);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4* w4 = reinterpret_cast<const float4*>(wts + (n0 + u) * LT_TAIL);
#pragma unroll
                for (int t4 = 0; t4 < LT_TAIL / 4; ++t4) {
                    const float4 w = w4[t4];
                    acc[t4 * 4 + 0] = fmaf(y[u], w.x, acc[t4 * 4 + 0]);
                    acc[t4 * 4 + 1] = fmaf(y[u], w.y, acc[t4 * 4 + 1]);
                    acc[t4 * 4 + 2] = fmaf(y[u], w.z, acc[t4 * 4 + 2]);
                    acc[t4 * 4 + 3] = fmaf(y[u], w.w, acc[t4 * 4 + 3]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < LT_TAIL; ++t) {
            if (t < kt) {
                float v = acc[t];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) out[row * ldo + t] = v;
            }
        }
    }
}

int launch_tail_dgrad(const float* dy, int lddy, int ny, const float* pre, int ldpre, int act, const float* wtail, int wt_sn,
                      int wt_st, int kt, int64_t rows, float* out, int ldo, cudaStream_t s) {
    if (rows <= 0 || kt <= 0) return MGB_OK;
    MGB_REQUIRE(kt <= LT_TAIL && (ny == 128 || ny == 256), "tail_dgrad: kt <= 16 and ny in {128, 256}");
    const int64_t want = ceil_div<int64_t>(rows, 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    tail_dgrad_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(dy, lddy, ny, pre, ldpre, act, wtail, wt_sn, wt_st, kt, rows, out, ldo);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
