// magnet_b200 — a whole MLP (models/backbones/mlp.py:9-28: Linear(128,128) + ReLU ... Linear(128,out)) in ONE kernel.
//
// The per-layer tensor-core Linear (linear_tc.cu) streams [rows,128] fp32 activations through HBM once per layer; the
// 4+1-layer MLPs of MAgNet (projector over every (query, t) row, the edge / node functions of the 10 InteractionNetworks)
// are then HBM-bound on traffic that never needs to leave the SM.  Here a CTA keeps TWO 128-row tiles of activations in
// shared memory (fp16 hi | lo images, 64 KB each) next to ONE layer's weight images (64 KB, re-loaded from L2 with a bulk
// async copy per layer) and walks the layers:
//   layer 0     : producers load x rows (fp32, optional ReLU on load) -> K-major images          (as linear_tc.cu)
//   every layer : D^T[n][row] = sum_k W_l[n][k] X_l[row][k]   (three fp16-split MMA terms, small terms first, TMEM)
//   hidden layer: epilogue (thread = channel n) adds bias, ReLU, splits into fp16 hi | lo and writes X_{l+1} IN PLACE
//                 as an MN-major image [n][row] — exactly the operand form the next layer's MMA reads (B MN-major)
//   last layer  : epilogue adds bias and stores y[row][n] (n < n_out), coalesced per row
// The two tiles ping-pong: while the epilogue of one runs, the MMAs of the other do.  Forward only (inference / rollout);
// same arithmetic as mgb_linear_tc_fwd with precision 3 (fp16 hi/lo split, 22 significant bits).
#include "internal.cuh"
#include "tc_common.cuh"

namespace mgb {

constexpr int MC_EPI_WARPS = 8, MC_PROD_WARPS = 8;
constexpr int MC_MMA_WARP = MC_EPI_WARPS, MC_PROD_WARP0 = MC_EPI_WARPS + 1;
constexpr int MC_THREADS = (MC_PROD_WARP0 + MC_PROD_WARPS) * 32;      // 544
constexpr size_t MLP_CHAIN_SMEM = 1024 + (size_t)2 * TILE_BYTES + (size_t)4 * TILE_BYTES + 256;

__global__ void __launch_bounds__(MC_THREADS, 1) mlp_chain_tc_kernel(const MlpChainArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* w_img = base;                                        // [hi|lo] of the current layer
    unsigned char* x_img = base + (size_t)2 * TILE_BYTES;               // [tile 0|1][hi|lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(x_img + (size_t)4 * TILE_BYTES);
    uint64_t* x_full = bars;          // [2] producers -> MMA (layer 0 operand written)
    uint64_t* x_empty = bars + 2;     // [2] last-layer epilogue -> producers (tile slot and accumulator free)
    uint64_t* t_full = bars + 4;      // [2] MMA -> epilogue (accumulator of the layer ready)
    uint64_t* x_ready = bars + 6;     // [2] hidden-layer epilogue -> MMA (next operand written, accumulator drained)
    uint64_t* w_bar = bars + 8;       // bulk copy of a layer's weights landed
    uint64_t* w_free = bars + 9;      // every MMA that reads the current weights has completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.rows, 128);
    const int64_t n_pairs = (n_tiles + 1) / 2;
    const int np = (int)((n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tile pairs of this CTA (>= 1)
    const int L = a.n_layers;

    if (tid == 0) {
        for (int t = 0; t < 2; ++t) {
            umma::mbar_init(&x_full[t], MC_PROD_WARPS * 32);
            umma::mbar_init(&x_empty[t], MC_EPI_WARPS * 32);
            umma::mbar_init(&t_full[t], 1);
            umma::mbar_init(&x_ready[t], MC_EPI_WARPS * 32);
        }
        umma::mbar_init(w_bar, 1);
        umma::mbar_init(w_free, 1);
        umma::fence_barrier_init();
    }
    if (warp == MC_MMA_WARP) umma::tmem_alloc(tmem_slot, 256);
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
                    const int64_t r0 = tile * 128;
                    const int nr = (int)((a.rows - r0) < 128 ? (a.rows - r0) : 128);
                    unsigned char* xrow = x_img + (size_t)t * 2 * TILE_BYTES + n * 128;
                    umma::mbar_wait(&t_full[t], tf[t] & 1);
                    ++tf[t];
                    umma::tc_fence_after();
#pragma unroll 1
                    for (int cb = 0; cb < 64; cb += 8) {
                        const int c0 = hf * 64 + cb;
                        float v[8];
                        umma::tmem_ld8(tmem + (uint32_t)(t * 128) + lane_base + c0, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] += bias;
                        if (!last) {
                            if (a.act == ACT_RELU) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                            } else if (a.act == ACT_SWISH) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = swish_tc<false>(v[i]);
                            }
                            if (fmaxf(fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))),
                                      fmaxf(fmaxf(fabsf(v[4]), fabsf(v[5])), fmaxf(fabsf(v[6]), fabsf(v[7])))) >= 32768.f && a.range_flag)
                                *a.range_flag = 1;
                            // X_{l+1}[row = c0 .. c0+7][k = n] as an MN-major image: row n of the image, 16 bytes
                            const uint32_t off = (uint32_t)(c0 >> 6) * (128u * 128u) + (uint32_t)((((c0 & 63) >> 3) ^ (n & 7)) << 4);
                            uint4 hi, lo;
                            split2_f16(v[0], v[1], hi.x, lo.x);
                            split2_f16(v[2], v[3], hi.y, lo.y);
                            split2_f16(v[4], v[5], hi.z, lo.z);
                            split2_f16(v[6], v[7], hi.w, lo.w);
                            *reinterpret_cast<uint4*>(xrow + off) = hi;
                            *reinterpret_cast<uint4*>(xrow + TILE_BYTES + off) = lo;
                        } else if (n < a.n_out) {
                            float* yo = a.y + (r0 + c0) * a.ldy + n;
                            const int lim = nr - c0;
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (i < lim) yo[(int64_t)i * a.ldy] = v[i];
                        }
                    }
                    if (!last) {
                        umma::fence_async_smem();
                        umma::tc_fence_before();
                        umma::mbar_arrive(&x_ready[t]);
                    } else {
                        umma::tc_fence_before();
                        umma::mbar_arrive(&x_empty[t]);
                    }
                }
            }
        }
    } else if (warp == MC_MMA_WARP) {
        // =========================== MMA issue + weight loads =======================================
        const uint32_t id_k = umma::idesc_f16(128, 128, 0, 0);       // layer 0: B K-major (rows of x)
        const uint32_t id_m = umma::idesc_f16(128, 128, 0, 1);       // layers >= 1: B MN-major ([n][row] images)
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
            for (int l = 0; l < L; ++l) {
                if (wl > 0) umma::mbar_wait(w_free, (wl - 1) & 1);          // the MMAs of the previous layer are done with w_img
                if (lane == 0) load_w2_image(w_img, (const unsigned char*)a.wimg + (size_t)l * 2 * TILE_BYTES, 2 * TILE_BYTES, w_bar);
                umma::mbar_wait(w_bar, wl & 1);
                ++wl;
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    if (pair * 2 + t >= n_tiles) continue;
                    if (l == 0) { umma::mbar_wait(&x_full[t], xf[t] & 1); ++xf[t]; }
                    else { umma::mbar_wait(&x_ready[t], xr[t] & 1); ++xr[t]; }
                    umma::tc_fence_after();
                    if (umma::elect_one()) {
                        const uint32_t d = tmem + (uint32_t)(t * 128);
                        const uint64_t xd = (l == 0 ? xk_d : xm_d) + (uint64_t)((uint32_t)t * 2 * TB);
#pragma unroll
                        for (int term = 0; term < 3; ++term) {      // small terms first: lo*hi, hi*lo, hi*hi
                            const uint64_t wa = w_d + (term == 0 ? TB : 0), xb = xd + (term == 1 ? TB : 0);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t koff_k = (uint32_t)((k >> 2) * (128 * 128 >> 4) + (k & 3) * 2), koff_m = (uint32_t)(k * 128);
                                umma::mma_bf16(d, wa + (uint64_t)koff_k, xb + (uint64_t)(l == 0 ? koff_k : koff_m), l == 0 ? id_k : id_m,
                                               (term | k) ? 1u : 0u);
                            }
                        }
                        umma::mma_commit(&t_full[t]);
                    }
                    __syncwarp();
                }
                if (umma::elect_one()) umma::mma_commit(w_free);
                __syncwarp();
            }
        }
    } else {
        // =========================== producers: 16 rows per warp (layer-0 operand) ==================
        const int pw = warp - MC_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        const float* src = a.x + lane * 4;
        uint32_t xe[2] = {0, 0};
#pragma unroll 1
        for (int it = 0; it < np; ++it) {
            const int64_t pair = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                const int64_t tile = pair * 2 + t;
                if (tile >= n_tiles) continue;
                const int64_t r0 = tile * 128 + pw * 16;
                float4 x[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int64_t row = r0 + r < a.rows ? r0 + r : a.rows - 1;
                    x[r] = *reinterpret_cast<const float4*>(src + row * a.ldx);
                }
                uint4 hl[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    float4 h = x[r];
                    if (a.in_act == ACT_RELU) { h.x = fmaxf(h.x, 0.f); h.y = fmaxf(h.y, 0.f); h.z = fmaxf(h.z, 0.f); h.w = fmaxf(h.w, 0.f); }
                    if (r0 + r >= a.rows) h = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (fmaxf(fmaxf(fabsf(h.x), fabsf(h.y)), fmaxf(fabsf(h.z), fabsf(h.w))) >= 32768.f && a.range_flag) *a.range_flag = 1;
                    split2_f16(h.x, h.y, hl[r].x, hl[r].z);
                    split2_f16(h.z, h.w, hl[r].y, hl[r].w);
                }
                umma::mbar_wait(&x_empty[t], (xe[t] & 1) ^ 1);
                ++xe[t];
                unsigned char* img = x_img + (size_t)t * 2 * TILE_BYTES;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                    *reinterpret_cast<uint2*>(img + off) = make_uint2(hl[r].x, hl[r].y);
                    *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = make_uint2(hl[r].z, hl[r].w);
                }
                umma::fence_async_smem();
                umma::mbar_arrive(&x_full[t]);
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == MC_MMA_WARP) umma::tmem_dealloc(tmem, 256);
}

int launch_mlp_chain_tc(const MlpChainArgs& a, cudaStream_t s) {
    MGB_REQUIRE(a.n_layers >= 1 && a.n_layers <= 8, "mlp_chain: 1..8 layers (got %d)", a.n_layers);
    MGB_REQUIRE(a.n_out >= 1 && a.n_out <= 128, "mlp_chain: 1 <= n_out <= 128 (got %d)", a.n_out);
    MGB_REQUIRE(a.rows >= 0 && a.rows < ((int64_t)1 << 31), "mlp_chain: row count out of range");
    MGB_REQUIRE(a.ldx % 4 == 0 && ((uintptr_t)a.x % 16) == 0, "mlp_chain: x rows must be 16-byte aligned");
    if (a.rows == 0) return MGB_OK;
    const int64_t pairs = (ceil_div<int64_t>(a.rows, 128) + 1) / 2;
    const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
    MGB_CUDA(cudaFuncSetAttribute(mlp_chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_CHAIN_SMEM));
    ProfScope prof(PROF_NODE_GEMM, s);
    mlp_chain_tc_kernel<<<grid, MC_THREADS, MLP_CHAIN_SMEM, s>>>(a);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
