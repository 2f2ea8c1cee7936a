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

namespace mgb {

constexpr int TCH = 128;            // hidden width
constexpr int TCE = 128;            // edge positions per tile (MMA N)
constexpr int TILE_BYTES = 128 * 256;   // one [128][128] bf16 image
constexpr int TC_STAGES = 2;
constexpr int TC_EPI_WARPS = 4, TC_PROD_WARPS = 8;
constexpr int TC_THREADS = (TC_EPI_WARPS + 1 + TC_PROD_WARPS) * 32;

template <bool FAST>
__device__ __forceinline__ float sigmoid_tc(float x) {
    if (FAST) {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
        return fmaf(0.5f, t, 0.5f);
    }
    return sigmoid_f(x);
}

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

__global__ void __launch_bounds__(TCH)
segment_fixup_tc_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ dstv, int64_t n_edges, int te,
                        const float* __restrict__ part_head, const float* __restrict__ part_tail, float* __restrict__ out,
                        int ld_out, int mean) {
    const int64_t t = (int64_t)blockIdx.x + 1;
    const int64_t e0 = t * te;
    if (e0 >= n_edges) return;
    const int node = dstv[e0];
    const int64_t s0 = rowptr[node], s1 = rowptr[node + 1];
    if (!(s0 < e0 && s1 <= e0 + te)) return;
    const int c = threadIdx.x;
    const int64_t t0 = s0 / te;
    float acc = part_tail[t0 * TCH + c];
    for (int64_t tt = t0 + 1; tt <= t; ++tt) acc += part_head[tt * TCH + c];
    out[(int64_t)node * ld_out + c] = mean ? acc / (float)(s1 - s0) : acc;
}

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
    return 1024 + (size_t)NSPLIT * TILE_BYTES * (1 + TC_STAGES) + 2 * TCE * sizeof(int) + 256;
}

template <int NSPLIT, bool FAST>
__global__ void __launch_bounds__(TC_THREADS, 1) gnn_edge_fwd_tc_kernel(const EdgeFwdTcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* w_img = base;                                         // NSPLIT images
    unsigned char* b_img = base + (size_t)NSPLIT * TILE_BYTES;            // [stage][split]
    int* epi_dst = reinterpret_cast<int*>(b_img + (size_t)TC_STAGES * NSPLIT * TILE_BYTES);   // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_dst + 2 * TCE);
    uint64_t* full = bars;                    // [stages] producers -> MMA
    uint64_t* empty = bars + TC_STAGES;       // [stages] MMA -> producers
    uint64_t* tfull = bars + 2 * TC_STAGES;   // [2] MMA -> epilogue
    uint64_t* tempty = bars + 2 * TC_STAGES + 2;   // [2] epilogue -> MMA
    uint64_t* wbar = bars + 2 * TC_STAGES + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_edges, TCE);

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            umma::mbar_init(&full[s], TC_PROD_WARPS * 32);
            umma::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&tfull[s], 1);
            umma::mbar_init(&tempty[s], TC_EPI_WARPS * 32);
        }
        umma::mbar_init(wbar, 1);
        umma::fence_barrier_init();
    }
    if (warp == TC_EPI_WARPS) umma::tmem_alloc(tmem_slot, 256);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < TC_EPI_WARPS) {
        // =========================== epilogue: thread = output channel n ===========================
        const int n = tid;
        const float bias = a.b2[n];
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int64_t e0 = tile * TCE;
            const int ne = (int)((a.n_edges - e0) < (int64_t)TCE ? (a.n_edges - e0) : (int64_t)TCE);
            int* dsts = epi_dst + acc * TCE;
            dsts[tid] = tid < ne ? a.dstv[e0 + tid] : -1;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            umma::mbar_wait(&tfull[acc], aph);
            umma::tc_fence_after();
            const int64_t e1 = e0 + ne;
            int cur = dsts[0];
            float sum = 0.f;
            auto flush = [&](int node, float v) {
                const int64_t s0 = a.rowptr[node], s1 = a.rowptr[node + 1];
                if (s0 >= e0 && s1 <= e1) a.agg[(int64_t)node * TCH + n] = v / (float)(s1 - s0);
                else if (s0 < e0) a.part_head[tile * TCH + n] = v;
                else a.part_tail[tile * TCH + n] = v;
            };
#pragma unroll 1
            for (int c0 = 0; c0 < TCE; c0 += 32) {
                float v[32];
                umma::tmem_ld32(tmem + (uint32_t)(acc * TCE) + ((uint32_t)(warp * 32) << 16) + c0, v);
                if (c0 + 32 >= TCE) {          // accumulator fully drained into registers: hand it back
                    umma::tc_fence_before();
                    umma::mbar_arrive(&tempty[acc]);
                }
                if (c0 < ne) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int e = c0 + i;
                        if (e < ne) {
                            const int d = dsts[e];
                            if (d != cur) {
                                flush(cur, sum);
                                cur = d;
                                sum = 0.f;
                            }
                            const float z = v[i] + bias;
                            sum += z * sigmoid_tc<FAST>(z);
                        }
                    }
                }
            }
            flush(cur, sum);
        }
    } else if (warp == TC_EPI_WARPS) {
        // =========================== MMA issue ====================================================
        if (lane == 0) {
            // W2 image(s): one bulk async copy (TMA engine, no tensor map needed for a pre-swizzled image)
            const uint32_t bytes = NSPLIT * TILE_BYTES;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(wbar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             umma::smem_u32(w_img)),
                         "l"(a.w2img), "r"(bytes), "r"(umma::smem_u32(wbar))
                         : "memory");
        }
        umma::mbar_wait(wbar, 0);
        const uint32_t idesc = umma::idesc_bf16(128, 128, 0, 0);
        const uint32_t w_s = umma::smem_u32(w_img);
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = it % TC_STAGES;
            const uint32_t ph = (it / TC_STAGES) & 1;
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            umma::mbar_wait(&tempty[acc], aph ^ 1);
            umma::mbar_wait(&full[s], ph);
            umma::tc_fence_after();
            if (lane == 0) {
                const uint32_t b_s = umma::smem_u32(b_img + (size_t)s * NSPLIT * TILE_BYTES);
                const uint32_t d = tmem + (uint32_t)(acc * TCE);
                uint32_t accum = 0;
#pragma unroll
                for (int term = 0; term < (NSPLIT == 1 ? 1 : 3); ++term) {
                    const int wa = term == 2 ? 1 : 0, hb = term == 1 ? 1 : 0;     // hi*hi, hi*lo, lo*hi
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        umma::mma_bf16(d, umma::desc_kmajor(w_s + wa * TILE_BYTES, k), umma::desc_kmajor(b_s + hb * TILE_BYTES, k),
                                       idesc, accum);
                        accum = 1;
                    }
                }
                umma::mma_commit(&empty[s]);
                umma::mma_commit(&tfull[acc]);
            }
            __syncwarp();
        }
    } else {
        // =========================== producers: 16 consecutive edge rows per warp ===================
        const int pw = warp - TC_EPI_WARPS - 1;
        constexpr int ROWS = TCE / TC_PROD_WARPS;   // 16
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = it % TC_STAGES;
            const uint32_t ph = (it / TC_STAGES) & 1;
            const int64_t e0 = tile * TCE + pw * ROWS;
            int my_d = -1, my_s = -1;
            if (lane < ROWS && e0 + lane < a.n_edges) {
                my_d = a.dstv[e0 + lane];
                my_s = a.srcv[e0 + lane];
            }
            umma::mbar_wait(&empty[s], ph ^ 1);
            unsigned char* img = b_img + (size_t)s * NSPLIT * TILE_BYTES;
            int prev_d = -2;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int r = 0; r < ROWS; ++r) {
                const int d = __shfl_sync(0xffffffffu, my_d, r);
                const int sidx = __shfl_sync(0xffffffffu, my_s, r);
                float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                if (d >= 0) {
                    if (d != prev_d) {
                        p = *reinterpret_cast<const float4*>(a.pq + (int64_t)d * (2 * TCH) + lane * 4);
                        prev_d = d;
                    }
                    const float4 q = *reinterpret_cast<const float4*>(a.pq + (int64_t)sidx * (2 * TCH) + TCH + lane * 4);
                    float z;
                    z = p.x + q.x; h.x = z * sigmoid_tc<FAST>(z);
                    z = p.y + q.y; h.y = z * sigmoid_tc<FAST>(z);
                    z = p.z + q.z; h.z = z * sigmoid_tc<FAST>(z);
                    z = p.w + q.w; h.w = z * sigmoid_tc<FAST>(z);
                }
                const uint32_t off = umma::tile_off(128, pw * ROWS + r, lane * 4);
                if (NSPLIT == 1) {
                    *reinterpret_cast<uint2*>(img + off) = make_uint2(umma::pack_bf16(h.x, h.y), umma::pack_bf16(h.z, h.w));
                } else {
                    __nv_bfloat16 hi[4], lo[4];
                    umma::split_bf16(h.x, hi[0], lo[0]);
                    umma::split_bf16(h.y, hi[1], lo[1]);
                    umma::split_bf16(h.z, hi[2], lo[2]);
                    umma::split_bf16(h.w, hi[3], lo[3]);
                    *reinterpret_cast<uint2*>(img + off) = *reinterpret_cast<uint2*>(hi);
                    *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = *reinterpret_cast<uint2*>(lo);
                }
            }
            umma::fence_async_smem();
            umma::mbar_arrive(&full[s]);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == TC_EPI_WARPS) umma::tmem_dealloc(tmem, 256);
}

size_t edge_fwd_tc_workspace(int64_t n_edges) {
    const int64_t tiles = ceil_div<int64_t>(n_edges > 0 ? n_edges : 1, TCE);
    return 2 * align_up((size_t)tiles * TCH * sizeof(float)) + 512;
}

// precision: 1 = bf16 hi/lo split (fp32 contract), 2 = plain bf16 + tanh.approx Swish
int launch_edge_fwd_tc(int precision, const float* pq, const int32_t* rowptr, const int32_t* dstv, const int32_t* srcv,
                       int64_t n_edges, const void* w2img, const float* b2, float* agg, void* ws_ptr, size_t ws_bytes,
                       cudaStream_t s) {
    if (n_edges <= 0) return MGB_OK;
    const int64_t tiles = ceil_div<int64_t>(n_edges, TCE);
    Workspace ws(ws_ptr, ws_bytes);
    float* part_head = ws.take<float>((size_t)tiles * TCH);
    float* part_tail = ws.take<float>((size_t)tiles * TCH);
    MGB_WS_CHECK(ws);
    EdgeFwdTcArgs a{pq, rowptr, dstv, srcv, n_edges, (const unsigned char*)w2img, b2, agg, part_head, part_tail};
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    {
        ProfScope prof(PROF_EDGE_FWD, s);
        if (precision == 2) {
            constexpr size_t smem = edge_fwd_tc_smem<1>();
            MGB_CUDA(cudaFuncSetAttribute(gnn_edge_fwd_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gnn_edge_fwd_tc_kernel<1, true><<<grid, TC_THREADS, smem, s>>>(a);
        } else {
            constexpr size_t smem = edge_fwd_tc_smem<2>();
            MGB_CUDA(cudaFuncSetAttribute(gnn_edge_fwd_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gnn_edge_fwd_tc_kernel<2, false><<<grid, TC_THREADS, smem, s>>>(a);
        }
    }
    MGB_LAUNCH_CHECK();
    if (tiles > 1) {
        segment_fixup_tc_kernel<<<(unsigned)(tiles - 1), TCH, 0, s>>>(rowptr, dstv, n_edges, TCE, part_head, part_tail, agg, TCH, 1);
        MGB_LAUNCH_CHECK();
    }
    return MGB_OK;
}

}  // namespace mgb
