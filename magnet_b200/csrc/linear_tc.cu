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

constexpr int LT_EPI_WARPS = 8, LT_PROD_WARPS = 8;
// whole warpgroups, so that registers can follow the work (setmaxnreg): warps 0-7 epilogue, 8 MMA issue, 9-11 idle,
// 12-19 producers; 640 threads x 96 registers at launch -> 112 for the epilogue, 104 for the producers, 40 for the rest
// (setmaxnreg.inc can only take what the CTA's own warps have released: (96 - 40) x 128 = 7168 >= 16 x 256 + 8 x 256)
constexpr int LT_MMA_WARP = LT_EPI_WARPS, LT_PROD_WARP0 = LT_EPI_WARPS + 4;
constexpr int LT_THREADS = (LT_PROD_WARP0 + LT_PROD_WARPS) * 32;      // 640
constexpr int LT_REGS_EPI = 112, LT_REGS_PROD = 104, LT_REGS_IDLE = 40;
constexpr int LT_TAIL = 16;

#ifdef MGB_TIMELINE
__device__ long long* g_lt_timeline = nullptr;      // [role 0..4][tile 0..7][event 0..3] of CTA 0 (tools/dev_lt_timeline.py); roles 3, 4: wgrad producer / MMA
#define LTTL(role, it_, ev) do { if (blockIdx.x == 0 && (it_) < 8 && (threadIdx.x & 31) == 0 && g_lt_timeline) g_lt_timeline[((role) * 8 + (it_)) * 4 + (ev)] = clock64(); } while (0)
int set_lt_timeline_buffer(long long* p) {
    return cudaMemcpyToSymbol(g_lt_timeline, &p, sizeof(p)) == cudaSuccess ? MGB_OK : MGB_ERR_CUDA;
}
#else
#define LTTL(role, it_, ev) do { } while (0)
#endif

// L2 prefetch of 16 rows x 512 bytes (64 lines of 128 bytes, two per lane): the row loads of the next tile and the 4-byte
// residual loads of the epilogue then hit L2 instead of paying the HBM latency in the middle of the pipeline
__device__ __forceinline__ void prefetch_rows16(const float* base, int64_t ld, int64_t r0, int64_t n_rows, int lane) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int line = j * 32 + lane;            // row = line / 4, 128-byte line inside the row = line % 4
        const int64_t row = r0 + (line >> 2);
        if (row < n_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + row * ld + (line & 3) * 32));
    }
}

// W [rows_total][ld] fp32, tile = W[r0:r0+128, c0:c0+128] (zero outside) -> swizzled bf16 images hi | lo
__global__ void pack_weight_tile_kernel(const float* __restrict__ W, int ld, int n_rows, int n_cols, int r0, int c0,
                                        unsigned char* __restrict__ img, int f16) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * 128) return;
    const int r = idx >> 7, c = idx & 127;
    const float v = (r0 + r < n_rows && c0 + c < n_cols) ? W[(int64_t)(r0 + r) * ld + c0 + c] : 0.f;
    const uint32_t off = umma::tile_off(128, r, c);
    if (f16) {
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        *reinterpret_cast<__half*>(img + off) = hi;
        *reinterpret_cast<__half*>(img + TILE_BYTES + off) = lo;
    } else {
        __nv_bfloat16 hi, lo;
        umma::split_bf16(v, hi, lo);
        *reinterpret_cast<__nv_bfloat16*>(img + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(img + TILE_BYTES + off) = lo;
    }
}

// The same 128 x 128 block as fp16 hi | lo pairs in the order the weight-loader warps of mlp_chain_tc.cu read it (A operand in
// tensor memory): 32-bit word j of row m = elements (2j, 2j+1); part (hi, lo) x chunk (j / 32) x w4 ((j % 32) / 4) x m x 4 words,
// so that thread m's eight 16-byte loads of a chunk are coalesced across the warp.  16384 words = 64 KB per block.
__global__ void pack_weight_tmem_kernel(const float* __restrict__ W, int ld, int n_rows, int n_cols, uint32_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * 64) return;
    const int m = idx >> 6, j = idx & 63;
    const float v0 = (m < n_rows && 2 * j < n_cols) ? W[(int64_t)m * ld + 2 * j] : 0.f;
    const float v1 = (m < n_rows && 2 * j + 1 < n_cols) ? W[(int64_t)m * ld + 2 * j + 1] : 0.f;
    uint32_t hi, lo;
    split2_f16(v0, v1, hi, lo);
    const int word = (((j >> 5) * 8 + ((j & 31) >> 2)) * 128 + m) * 4 + (j & 3);
    out[word] = hi;
    out[8192 + word] = lo;
}

int pack_weight_tmem(const float* W, int ld, int n_rows, int n_cols, void* out, cudaStream_t s) {
    pack_weight_tmem_kernel<<<32, 256, 0, s>>>(W, ld, n_rows, n_cols, (uint32_t*)out);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

int pack_weight_tile(const float* W, int ld, int n_rows, int n_cols, int r0, int c0, void* img, cudaStream_t s, int f16) {
    pack_weight_tile_kernel<<<64, 256, 0, s>>>(W, ld, n_rows, n_cols, r0, c0, (unsigned char*)img, f16);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// 2 weight tiles (always hi|lo) + one B stage + 2 tail tiles + the tail weights; with a single weight tile the second
// weight slot serves as a second B stage
constexpr size_t LINEAR_TC_SMEM = 1024 + (size_t)4 * TILE_BYTES + (size_t)2 * TILE_BYTES + 4 * 128 * LT_TAIL * sizeof(float) + 256;

// All role loops are kept small on purpose: the three roles of a CTA run at the same time and share the SM's
// 32 KB instruction cache (a fully unrolled version of this kernel was 100 KB of SASS and fetch-bound).
// F16: operands split into two fp16 values instead of two bf16 values (22 instead of 16 significant bits at the same
// cost; needs |x| < 65504 and O(1) data: used for the forward Linears of the MLPs, whose inputs are normalised)
template <int NSPLIT, bool FAST, bool F16 = false>
__global__ void __launch_bounds__(LT_THREADS, 1) linear_tc_kernel(const LinTcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* w_img = base;                                        // [2 tiles][hi|lo]
    unsigned char* b_img = w_img + (size_t)4 * TILE_BYTES;               // B stage 0 (NSPLIT = 1: stages 0 and 1)
    float* tails = reinterpret_cast<float*>(b_img + (size_t)2 * TILE_BYTES);   // [2][128][LT_TAIL]
    float* wts = tails + 2 * 128 * LT_TAIL;                                    // [2 output blocks][LT_TAIL][128] tail weights
    uint64_t* bars = reinterpret_cast<uint64_t*>(wts + 2 * 128 * LT_TAIL);
    uint64_t* full = bars;            // [2]
    uint64_t* empty = bars + 2;       // [2]
    uint64_t* tfull = bars + 4;       // [2] accumulators ready
    uint64_t* tempty = bars + 6;      // [2] accumulators (and tail tile) drained
    uint64_t* tailfull = bars + 8;    // [2] tail tile written
    uint64_t* wbar = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.rows, 128);
    const int nt = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int n_wtiles = a.nm * a.nk;
    const int bstages = (NSPLIT == 1 || n_wtiles == 1) ? 2 : 1;
    // stage s of the B operand
    auto b_stage = [&](int s) -> unsigned char* {
        if (NSPLIT == 1) return b_img + (size_t)s * TILE_BYTES;
        return s == 0 ? b_img : w_img + (size_t)2 * TILE_BYTES;
    };

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
        umma::reg_inc<LT_REGS_EPI>();
        // =========================== epilogue: thread = output channel within the block; warps 0-3 rows 0-63 of the
        // tile, warps 4-7 rows 64-127 ====================================================================
        // Sixteen rows per step, software-pipelined: the TMEM load and the residual loads of the next step are in flight
        // while this one is processed (two register sets alternate), and the first residual loads of a tile go out
        // before its accumulator is waited for.  Addressing is one walking pointer per stream (y, y_pre, residual) and
        // 32-bit row limits: with 64-bit index arithmetic per element the 4-byte accesses cost ~9 instructions each and
        // the epilogue, not HBM, bound the kernel (ncu: 49 warp-instructions per output element before, round 2).
        const int n = tid & 127;
        const int hf = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int ldy = a.ldy, ldyp = a.ldyp, ldr = a.ldr, kt = a.kt, act = a.act;
        float bias_m[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int col = m * 128 + n;
            const bool col_ok = m < a.nm && (a.n_out == 0 || col < a.n_out);
            bias_m[m] = (a.bias && col_ok) ? a.bias[col] : 0.f;
            // the tail weights of this thread's output column: wts[m][t][n] (both halves write the same values; a thread
            // only ever reads what it wrote itself)
            if (kt > 0) {
                for (int t = 0; t < LT_TAIL; ++t)
                    wts[(m * LT_TAIL + t) * 128 + n] = (col_ok && t < kt) ? a.wtail[(int64_t)col * a.wt_sn + (int64_t)t * a.wt_st] : 0.f;
            }
        }
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 128;
            const int nr = (int)((a.rows - r0) < 128 ? (a.rows - r0) : 128);
            const int rlim = nr - hf * 64;           // valid rows among the 64 of this half
            const int64_t rb = r0 + hf * 64;
            const float* tl = tails + acc * 128 * LT_TAIL + hf * 64 * LT_TAIL;
            bool waited = false;
#pragma unroll 1
            for (int m = 0; m < a.nm; ++m) {
                const int col = m * 128 + n;
                const bool col_ok = a.n_out == 0 || col < a.n_out;
                const float bias = m ? bias_m[1] : bias_m[0];
                const float* wt = wts + m * LT_TAIL * 128 + n;
                const bool last_m = m == a.nm - 1;
                const int vlim = col_ok ? rlim : 0;      // rows this thread stores
                const bool has_res = a.residual && col_ok && (a.res_blocks == 0 || ((a.res_blocks >> m) & 1));
                const float* rp = has_res ? a.residual + rb * ldr + col : nullptr;      // walks one step ahead of yo
                float* yo = a.y + rb * ldy + col;
                float* yp = a.y_pre ? a.y_pre + rb * ldyp + col : nullptr;
                const uint32_t tacc = tmem + (uint32_t)((acc * 2 + m) * 128) + lane_base + (uint32_t)(hf * 64);
                // residual values of rows c .. c+15 (4-byte loads, one coalesced 128-byte line per warp and row)
                auto load_res = [&](float (&dst)[16], int c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        dst[i] = 0.f;
                        if (c + i < vlim) dst[i] = *rp;
                        rp += ldr;
                    }
                };
                auto step = [&](int c, const uint32_t (&raw)[16], const float (&res)[16]) {
                    if (c >= vlim) return;
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
                    if (kt > 0) {
#pragma unroll 1
                        for (int t4 = 0; t4 * 4 < kt; ++t4) {
                            const float w0 = wt[(t4 * 4 + 0) * 128], w1 = wt[(t4 * 4 + 1) * 128], w2 = wt[(t4 * 4 + 2) * 128], w3 = wt[(t4 * 4 + 3) * 128];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float4 x = *reinterpret_cast<const float4*>(tl + (c + i) * LT_TAIL + t4 * 4);
                                v[i] = fmaf(x.x, w0, v[i]);
                                v[i] = fmaf(x.y, w1, v[i]);
                                v[i] = fmaf(x.z, w2, v[i]);
                                v[i] = fmaf(x.w, w3, v[i]);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += bias;
                    const int lim = vlim - c;
                    if (yp) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (i < lim) *yp = v[i];
                            yp += ldyp;
                        }
                    }
                    if (act == ACT_SWISH) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = swish_tc<FAST>(v[i]);
                    } else if (act == ACT_RELU) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (has_res) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += res[i];
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (i < lim) *yo = v[i];
                        yo += ldy;
                    }
                };
                uint32_t va[16], vb[16];
                float ra[16], rc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { ra[i] = 0.f; rc[i] = 0.f; }
                if (has_res) load_res(ra, 0);
                if (!waited) {
                    if (warp == 0) LTTL(2, it, 0);
                    if (kt > 0) umma::mbar_wait(&tailfull[acc], aph);
                    umma::mbar_wait(&tfull[acc], aph);
                    umma::tc_fence_after();
                    waited = true;
                    if (warp == 0) LTTL(2, it, 1);
                }
                umma::tmem_ld16_issue(tacc, va);
#pragma unroll 1
                for (int c = 0; c < 64; c += 32) {
                    if (has_res) load_res(rc, c + 16);
                    umma::tmem_ld_wait16(va);
                    if (warp == 0 && c == 0 && m == 0) LTTL(2, it, 2);
                    umma::tmem_ld16_issue(tacc + (uint32_t)(c + 16), vb);
                    step(c, va, ra);
                    if (has_res && c + 32 < 64) load_res(ra, c + 32);
                    umma::tmem_ld_wait16(vb);
                    if (c + 32 < 64) {
                        umma::tmem_ld16_issue(tacc + (uint32_t)(c + 32), va);
                    } else if (last_m && kt == 0) {      // this thread's part of the accumulators is in registers: hand them back
                        umma::tc_fence_before();
                        umma::mbar_arrive(&tempty[acc]);
                    }
                    step(c + 16, vb, rc);
                }
            }
            if (kt > 0) {          // the tail tile is read until the end: release accumulators and tail together
                umma::tc_fence_before();
                umma::mbar_arrive(&tempty[acc]);
            }
            if (warp == 0) LTTL(2, it, 3);
        }
    } else if (warp < LT_PROD_WARP0 && warp != LT_MMA_WARP) {
        umma::reg_dec<LT_REGS_IDLE>();          // padding warps of the MMA warpgroup
    } else if (warp == LT_MMA_WARP) {
        umma::reg_dec<LT_REGS_IDLE>();
        // =========================== MMA issue ====================================================
        if (lane == 0) load_w2_image(w_img, (const unsigned char*)a.wimg, (uint32_t)(n_wtiles * 2 * TILE_BYTES), wbar);   // global images always hold hi|lo
        umma::mbar_wait(wbar, 0);
        const uint32_t idesc = F16 ? umma::idesc_f16(128, 128, a.a_trans, 0) : umma::idesc_bf16(128, 128, a.a_trans, 0);
        // descriptors of k-step 0; a k-step only moves the start-address field (bytes >> 4)
        const uint64_t w_d = a.a_trans ? umma::desc_sw128(umma::smem_u32(w_img), 128 * 128, 1024) : umma::desc_sw128(umma::smem_u32(w_img), 16, 1024);
        const uint64_t b_d0 = umma::desc_sw128(umma::smem_u32(b_stage(0)), 16, 1024);
        const uint64_t b_d1 = umma::desc_sw128(umma::smem_u32(b_stage(1)), 16, 1024);
        constexpr uint32_t TB = TILE_BYTES >> 4;
        int sc = 0;
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            umma::mbar_wait(&tempty[acc], aph ^ 1);
#pragma unroll 1
            for (int kc = 0; kc < a.nk; ++kc, ++sc) {
                const int s = sc % bstages;
                umma::mbar_wait(&full[s], (sc / bstages) & 1);
                umma::tc_fence_after();
                LTTL(1, it, kc == 0 ? 0 : 2);
                if (umma::elect_one()) {
                    const uint64_t bd = s ? b_d1 : b_d0;
#pragma unroll 1
                    for (int m = 0; m < a.nm; ++m) {
                        const uint32_t d = tmem + (uint32_t)((acc * 2 + m) * 128);
                        const uint64_t wd = w_d + (uint64_t)((uint32_t)a.tile_of[m][kc] * 2 * TB);
#pragma unroll
                        for (int term = 0; term < (NSPLIT == 1 ? 1 : 3); ++term) {
                            // bf16: hi*hi, hi*lo, lo*hi.  fp16: the small terms first (lo*hi, hi*lo, hi*hi), so that the
                            // tensor core's truncating fp32 accumulation only rounds the eight big steps at full magnitude
                            const bool w_lo = F16 ? term == 0 : term == 2, b_lo = term == 1;
                            const uint64_t wa = wd + (w_lo ? TB : 0), bb = bd + (b_lo ? TB : 0);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t koff_k = (uint32_t)((k >> 2) * (128 * 128 >> 4) + (k & 3) * 2), koff_m = (uint32_t)(k * 128);
                                umma::mma_bf16(d, wa + (uint64_t)(a.a_trans ? koff_m : koff_k), bb + (uint64_t)koff_k, idesc,
                                               (kc | term | k) ? 1u : 0u);
                            }
                        }
                    }
                    umma::mma_commit(&empty[s]);
                    if (kc == a.nk - 1) umma::mma_commit(&tfull[acc]);
                }
                __syncwarp();
                LTTL(1, it, kc == 0 ? 1 : 3);
            }
        }
    } else {
        umma::reg_inc<LT_REGS_PROD>();
        // =========================== producers: 16 rows per warp ===================================
        // x rows -> (optional act'(pre) / act) -> bf16 hi/lo in place in the registers the rows came in, all before the
        // stage is waited for; only the stores follow its release.  Rows are reached through one walking pointer that
        // stops at the warp's last valid row (rows past the end repeat it and are zeroed afterwards).
        const int pw = warp - LT_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        // this lane's slot of the tail tile: t = lane & 15 for every element e = j*32 + lane it writes (row j*2 + lane / 16)
        const float* tail_p = nullptr;
        int tail_ld = 0;
        if (a.kt > 0 && (lane & 15) < a.kt) {
            int tt = lane & 15, seg = 0;
            while (tt >= a.tk[seg]) { tt -= a.tk[seg]; ++seg; }
            tail_p = a.tsrc[seg] + tt;
            tail_ld = a.tld[seg];
        }
        int sc = 0;
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 128 + pw * 16;
            const int64_t left = a.rows - r0;
            const int nv = left >= 16 ? 16 : (left > 0 ? (int)left : 0);       // valid rows of this warp
            const int64_t rfirst = nv > 0 ? r0 : a.rows - 1;
            {
                const int64_t rn = r0 + (int64_t)gridDim.x * 128;      // this warp's rows of the CTA's next tile
                for (int kc = 0; kc < a.nk; ++kc) prefetch_rows16(a.src[kc], a.ld[kc], rn, a.rows, lane);
                if (a.pre) prefetch_rows16(a.pre, a.ldpre, rn, a.rows, lane);
                if (a.residual) prefetch_rows16(a.residual, a.ldr, r0, a.rows, lane);      // read by the epilogue of this tile
            }
            if (a.kt > 0) {
                umma::mbar_wait(&tempty[acc], aph ^ 1);
                float* tl = tails + acc * 128 * LT_TAIL + pw * 16 * LT_TAIL + lane;
                // 16 rows x 16 tail slots = 256 values, 8 per lane
                const float* tp = tail_p ? tail_p + (r0 + (lane >> 4)) * tail_ld : nullptr;
                const int rr = lane >> 4;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float v = 0.f;
                    if (tp && j * 2 + rr < nv) v = *tp;
                    if (tp) tp += 2 * tail_ld;
                    tl[j * 32] = v;
                }
                umma::mbar_arrive(&tailfull[acc]);
            }
#pragma unroll 1
            for (int kc = 0; kc < a.nk; ++kc, ++sc) {
                const int s = sc % bstages;
                const int ld = a.ld[kc];
                float4 x[16];
                {
                    const float* p = a.src[kc] + rfirst * ld + lane * 4;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        x[r] = *reinterpret_cast<const float4*>(p);
                        if (r + 1 < nv) p += ld;
                    }
                }
                if (kc == 0 && a.pre) {
                    // pre-activation rows in two halves of eight, the first requested together with the x rows: two exposed
                    // memory latencies per tile instead of five
                    const float* p = a.pre + rfirst * a.ldpre + lane * 4;
                    const int ldp = a.ldpre;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 q[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            q[r] = *reinterpret_cast<const float4*>(p);
                            if (h * 8 + r + 1 < nv) p += ldp;
                        }
                        if (a.pre_act == ACT_SWISH) {
#pragma unroll
                            for (int r = 0; r < 8; ++r) {
                                x[h * 8 + r].x *= swish_grad_tc<FAST>(q[r].x);
                                x[h * 8 + r].y *= swish_grad_tc<FAST>(q[r].y);
                                x[h * 8 + r].z *= swish_grad_tc<FAST>(q[r].z);
                                x[h * 8 + r].w *= swish_grad_tc<FAST>(q[r].w);
                            }
                        } else if (a.pre_act == ACT_RELU) {
#pragma unroll
                            for (int r = 0; r < 8; ++r) {
                                x[h * 8 + r].x = q[r].x > 0.f ? x[h * 8 + r].x : 0.f;
                                x[h * 8 + r].y = q[r].y > 0.f ? x[h * 8 + r].y : 0.f;
                                x[h * 8 + r].z = q[r].z > 0.f ? x[h * 8 + r].z : 0.f;
                                x[h * 8 + r].w = q[r].w > 0.f ? x[h * 8 + r].w : 0.f;
                            }
                        }
                    }
                }
                if (kc == 0 && a.self_act == ACT_SWISH) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        x[r].x = swish_tc<FAST>(x[r].x); x[r].y = swish_tc<FAST>(x[r].y);
                        x[r].z = swish_tc<FAST>(x[r].z); x[r].w = swish_tc<FAST>(x[r].w);
                    }
                } else if (kc == 0 && a.self_act == ACT_RELU) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        x[r].x = fmaxf(x[r].x, 0.f); x[r].y = fmaxf(x[r].y, 0.f);
                        x[r].z = fmaxf(x[r].z, 0.f); x[r].w = fmaxf(x[r].w, 0.f);
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
                    const float4 h = x[r];
                    if (NSPLIT == 1) {
                        hl[r].x = umma::pack_bf16(h.x, h.y);
                        hl[r].y = umma::pack_bf16(h.z, h.w);
                    } else {
                        if (F16) {
                            // fp16 tops out at 65504: report values beyond half of that instead of producing infinities
                            if (fmaxf(fmaxf(fabsf(h.x), fabsf(h.y)), fmaxf(fabsf(h.z), fabsf(h.w))) >= 32768.f && a.range_flag) *a.range_flag = 1;
                            split2_f16(h.x, h.y, hl[r].x, hl[r].z);
                            split2_f16(h.z, h.w, hl[r].y, hl[r].w);
                        } else {
                            split2_bf16(h.x, h.y, hl[r].x, hl[r].z);
                            split2_bf16(h.z, h.w, hl[r].y, hl[r].w);
                        }
                    }
                }
                if (pw == 0 && kc == 0) LTTL(0, it, 0);
                umma::mbar_wait(&empty[s], ((sc / bstages) & 1) ^ 1);
                if (pw == 0 && kc == 0) LTTL(0, it, 1);
                unsigned char* img = b_stage(s);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                    *reinterpret_cast<uint2*>(img + off) = make_uint2(hl[r].x, hl[r].y);
                    if (NSPLIT == 2) *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = make_uint2(hl[r].z, hl[r].w);
                }
                umma::fence_async_smem();
                umma::mbar_arrive(&full[s]);
                if (pw == 0) LTTL(0, it, kc == 0 ? 2 : 3);
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
    } else if (precision == 3) {
        MGB_CUDA(cudaFuncSetAttribute(linear_tc_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LINEAR_TC_SMEM));
        linear_tc_kernel<2, false, true><<<grid, LT_THREADS, LINEAR_TC_SMEM, s>>>(a);
    } else {
        MGB_CUDA(cudaFuncSetAttribute(linear_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LINEAR_TC_SMEM));
        linear_tc_kernel<2, false><<<grid, LT_THREADS, LINEAR_TC_SMEM, s>>>(a);
    }
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ==================================================================================================
// Weight gradients on the tensor cores, all sources of one Linear in ONE launch:
//   dW[y][x][n][k] (+)= sum_rows Y'_y[row][n] X_x[row][k]
//   Y'_y = dy_y * act'(y_pre) (ny <= 2 tiles of 128 output channels),
//   X_x  = up to two 128-column sources (optionally act(x)) and a tail tile that packs the small-K columns (u, pos,
//          variables).  The bias gradient (column sums of Y') is accumulated by the producers on the way, in fp32.
// Both operands are row tiles [row][col] used as MN-major operands (K = rows).  The ny*(nx+1) <= 4 accumulators live in
// TMEM across the CTA's row tiles and are drained into the CTA's fp32 partial every WG_GROUP tiles with round-to-nearest
// adds (the tensor core's own accumulation truncates; see gnn_edge_tc.cu).
// Shared memory: Y' tiles 2 x 64 KB, one X stage 64 KB (with ny = 1 the second Y' slot is a second X stage).
// ==================================================================================================
constexpr size_t WGRAD_TC_SMEM = 1024 + (size_t)6 * TILE_BYTES + 256;
constexpr int WG_GROUP = 8;

template <int NSPLIT, bool FAST>
__global__ void __launch_bounds__(LT_THREADS, 1) wgrad_tc_kernel(const WgradTcArgs a) {
    constexpr int GROUP = FAST ? (1 << 30) : WG_GROUP;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = umma::smem_u32(smem_raw);
    unsigned char* base = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* y_img = base;                                  // [2][hi|lo]  Y'[row][n]
    unsigned char* x_img = base + (size_t)4 * TILE_BYTES;         // [hi|lo]     X[row][k]
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)6 * TILE_BYTES);
    uint64_t* y_full = bars;        // producers -> MMA: all Y' tiles of the row tile written
    uint64_t* y_empty = bars + 1;   // MMA -> producers: every MMA of the row tile has completed
    uint64_t* x_full = bars + 2;    // [2]
    uint64_t* x_empty = bars + 4;   // [2]
    uint64_t* d_full = bars + 6;    // MMA -> epilogue: a group of row tiles is complete
    uint64_t* d_empty = bars + 7;   // epilogue -> MMA: accumulators drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = ceil_div<int64_t>(a.rows, 128);
    const int nt = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int nxt = a.nx + (a.tail ? 1 : 0);      // X tiles per row tile
    const int nacc = a.ny * nxt;
    const int xstages = a.ny == 1 ? 2 : 1;
    auto x_stage = [&](int s) -> unsigned char* { return s == 0 ? x_img : y_img + (size_t)2 * TILE_BYTES; };
    if (tid == 0) {
        umma::mbar_init(y_full, LT_PROD_WARPS * 32);
        umma::mbar_init(y_empty, 1);
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&x_full[s], LT_PROD_WARPS * 32);
            umma::mbar_init(&x_empty[s], 1);
        }
        umma::mbar_init(d_full, 1);
        umma::mbar_init(d_empty, LT_EPI_WARPS * 32);
        umma::fence_barrier_init();
    }
    if (warp == LT_MMA_WARP) umma::tmem_alloc(tmem_slot, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < LT_EPI_WARPS) {
        umma::reg_dec<LT_REGS_IDLE + 24>();      // 64: with 40 for the MMA warpgroup this leaves 152 for the producers (640 x 96 in the pool)
        // =========================== drain: thread = output channel n; warps 0-3 columns 0-63, warps 4-7 columns 64-127
        const int n = tid & 127, hf = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int ngroups = (nt + GROUP - 1) / GROUP;
#pragma unroll 1
        for (int grp = 0; grp < ngroups; ++grp) {
            umma::mbar_wait_relaxed<500>(d_full, grp & 1);       // a group is eight row tiles away: do not poll in the producers' issue slots
            umma::tc_fence_after();
#pragma unroll 1
            for (int ac = 0; ac < nacc; ++ac) {
                float* out = a.partial + (((int64_t)blockIdx.x * nacc + ac) * 128 + n) * 128 + hf * 64;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    float4* o = reinterpret_cast<float4*>(out + c0);
                    float4 old[4];       // requested before the accumulator columns: the L2 latency overlaps the TMEM round trip
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) old[q4] = grp > 0 ? o[q4] : make_float4(0.f, 0.f, 0.f, 0.f);
                    float v[16];
                    umma::tmem_ld16(tmem + (uint32_t)(ac * 128 + hf * 64) + lane_base + c0, v);
                    if (ac == nacc - 1 && c0 + 16 >= 64) {
                        umma::tc_fence_before();
                        umma::mbar_arrive(d_empty);
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        float4 w = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
                        w.x += old[q4].x; w.y += old[q4].y; w.z += old[q4].z; w.w += old[q4].w;
                        o[q4] = w;
                    }
                }
            }
        }
    } else if (warp < LT_PROD_WARP0 && warp != LT_MMA_WARP) {
        umma::reg_dec<LT_REGS_IDLE>();          // padding warps of the MMA warpgroup
    } else if (warp == LT_MMA_WARP) {
        umma::reg_dec<LT_REGS_IDLE>();
        const uint32_t idesc = umma::idesc_bf16(128, 128, 1, 1);
        const uint64_t y_d = umma::desc_sw128(umma::smem_u32(y_img), 128 * 128, 1024);
        const uint64_t x_d0 = umma::desc_sw128(umma::smem_u32(x_stage(0)), 128 * 128, 1024);
        const uint64_t x_d1 = umma::desc_sw128(umma::smem_u32(x_stage(1)), 128 * 128, 1024);
        constexpr uint32_t TB = TILE_BYTES >> 4;
        int sc = 0;
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int grp = it / GROUP;
            const bool first_in_group = (it % GROUP) == 0;
            const bool last = it == nt - 1;
            if (first_in_group) umma::mbar_wait(d_empty, (grp & 1) ^ 1);
            umma::mbar_wait(y_full, it & 1);
            LTTL(4, it, 0);
#pragma unroll 1
            for (int xi = 0; xi < nxt; ++xi, ++sc) {
                const int s = sc % xstages;
                umma::mbar_wait(&x_full[s], (sc / xstages) & 1);
                umma::tc_fence_after();
                LTTL(4, it, xi == 0 ? 1 : 2);
                if (umma::elect_one()) {
                    const uint64_t xd = s ? x_d1 : x_d0;
#pragma unroll 1
                    for (int yi = 0; yi < a.ny; ++yi) {
                        const uint32_t d = tmem + (uint32_t)((yi * nxt + xi) * 128);
                        const uint64_t yd = y_d + (uint64_t)((uint32_t)yi * 2 * TB);
#pragma unroll
                        for (int term = 0; term < (NSPLIT == 1 ? 1 : 3); ++term) {
                            const uint64_t ya = yd + (term == 2 ? TB : 0), xb = xd + (term == 1 ? TB : 0);
#pragma unroll
                            for (int k = 0; k < 8; ++k)     // K = rows: 16 rows = 2048 bytes per step
                                umma::mma_bf16(d, ya + (uint64_t)(k * 128), xb + (uint64_t)(k * 128), idesc,
                                               (term | k) ? 1u : (first_in_group ? 0u : 1u));
                        }
                    }
                    umma::mma_commit(&x_empty[s]);
                    if (xi == nxt - 1) {
                        umma::mma_commit(y_empty);
                        if ((it % GROUP) == GROUP - 1 || last) umma::mma_commit(d_full);
                    }
                }
                __syncwarp();
                if (xi == nxt - 1) LTTL(4, it, 3);
            }
        }
    } else {
        umma::reg_inc<152>();
        // =========================== producers: 16 rows per warp ===================================
        const int pw = warp - LT_PROD_WARP0;
        const uint32_t lane_blk = (uint32_t)(lane >> 4) * (128u * 128u) + (uint32_t)(lane & 1) * 8u;
        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
        // 16 rows of one operand tile: converted in place in the registers the rows came in; only the stores follow the wait
        // 16 rows of one operand tile: rows past the end are zeroed, the split happens in the registers the rows came in
        auto convert = [&](float4 (&x)[16], int nv, uint4 (&hl)[16]) {
            if (nv < 16) {        // the last tile of the problem only
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    if (r >= nv) x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float4 h = x[r];
                if (NSPLIT == 1) {
                    hl[r].x = umma::pack_bf16(h.x, h.y);
                    hl[r].y = umma::pack_bf16(h.z, h.w);
                } else {
                    split2_bf16(h.x, h.y, hl[r].x, hl[r].z);
                    split2_bf16(h.z, h.w, hl[r].y, hl[r].w);
                }
            }
        };
        auto store = [&](const uint4 (&hl)[16], unsigned char* img) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const uint32_t off = lane_blk + (uint32_t)(pw * 16 + r) * 128u + ((lane_chunk ^ (uint32_t)(r & 7)) << 4);
                *reinterpret_cast<uint2*>(img + off) = make_uint2(hl[r].x, hl[r].y);
                if (NSPLIT == 2) *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = make_uint2(hl[r].z, hl[r].w);
            }
            umma::fence_async_smem();
        };
        // 16 rows x this lane's four columns through one walking pointer that stops at the warp's last valid row
        auto load_rows = [&](float4 (&x)[16], const float* p, int ld, int nv) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                x[r] = *reinterpret_cast<const float4*>(p);
                if (r + 1 < nv) p += ld;
            }
        };
        // this lane's four columns of the tail tile: source pointer (column already applied) and row stride, or a constant
        const float* tp[4];
        int tl[4];
        float tone[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = lane * 4 + u;
            tp[u] = nullptr; tl[u] = 0; tone[u] = 0.f;
            if (t < a.kt) {
                int tt = t, seg = 0;
                while (tt >= a.tk[seg]) { tt -= a.tk[seg]; ++seg; }
                tp[u] = a.tsrc[seg] + tt; tl[u] = a.tld[seg];
            }
        }
        // X operand xi of a row tile: a 128-column source, or the tail tile (columns [0, kt) = the small-K sources, the rest 0;
        // lanes >= 4 hold constants only)
        auto load_x = [&](float4 (&x)[16], int xi, int64_t rfirst, int nv) {
            if (xi < a.nx) {
                load_rows(x, a.x[xi] + rfirst * a.ldx[xi] + lane * 4, a.ldx[xi], nv);
            } else {
                const float* q0 = tp[0] ? tp[0] + rfirst * tl[0] : nullptr;
                const float* q1 = tp[1] ? tp[1] + rfirst * tl[1] : nullptr;
                const float* q2 = tp[2] ? tp[2] + rfirst * tl[2] : nullptr;
                const float* q3 = tp[3] ? tp[3] + rfirst * tl[3] : nullptr;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    x[r].x = q0 ? *q0 : tone[0];
                    x[r].y = q1 ? *q1 : tone[1];
                    x[r].z = q2 ? *q2 : tone[2];
                    x[r].w = q3 ? *q3 : tone[3];
                    if (r + 1 < nv) {
                        if (q0) q0 += tl[0];
                        if (q1) q1 += tl[1];
                        if (q2) q2 += tl[2];
                        if (q3) q3 += tl[3];
                    }
                }
            }
        };
        float4 bsum[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};     // column sums of Y' (this warp's rows)
        int sc = 0;
        // (measured and dropped: requesting the rows of the NEXT operand before the stage of the current one is waited for --
        // the two live row sets spill at 168 registers and the weight-gradient launches got 19 % slower)
#pragma unroll 1
        for (int it = 0; it < nt; ++it) {
            const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 128 + pw * 16;
            const int64_t left = a.rows - r0;
            const int nv = left >= 16 ? 16 : (left > 0 ? (int)left : 0);       // valid rows of this warp
            const int64_t rfirst = nv > 0 ? r0 : a.rows - 1;
            {
                const int64_t rn = r0 + (int64_t)gridDim.x * 128;      // this warp's rows of the CTA's next tile
                for (int yi = 0; yi < a.ny; ++yi) {
                    prefetch_rows16(a.dy + yi * 128, a.lddy, rn, a.rows, lane);
                    if (a.y_pre) prefetch_rows16(a.y_pre + yi * 128, a.ldyp, rn, a.rows, lane);
                }
                for (int xi = 0; xi < a.nx; ++xi) prefetch_rows16(a.x[xi], a.ldx[xi], rn, a.rows, lane);
            }
            float4 x[16];
            uint4 hl[16];
#pragma unroll 1
            for (int yi = 0; yi < a.ny; ++yi) {
                load_rows(x, a.dy + rfirst * a.lddy + yi * 128 + lane * 4, a.lddy, nv);
                if (a.y_pre) {
                    // the pre-activation rows are requested together with the dy rows: one exposed memory latency per
                    // Y' tile instead of five (timeline: ~2 k cycles each under load)
                    float4 q[16];
                    load_rows(q, a.y_pre + rfirst * a.ldyp + yi * 128 + lane * 4, a.ldyp, nv);
                    if (a.y_act == ACT_SWISH) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            x[r].x *= swish_grad_tc<FAST>(q[r].x);
                            x[r].y *= swish_grad_tc<FAST>(q[r].y);
                            x[r].z *= swish_grad_tc<FAST>(q[r].z);
                            x[r].w *= swish_grad_tc<FAST>(q[r].w);
                        }
                    } else if (a.y_act == ACT_RELU) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            x[r].x = q[r].x > 0.f ? x[r].x : 0.f;
                            x[r].y = q[r].y > 0.f ? x[r].y : 0.f;
                            x[r].z = q[r].z > 0.f ? x[r].z : 0.f;
                            x[r].w = q[r].w > 0.f ? x[r].w : 0.f;
                        }
                    }
                }
                {
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        if (r < nv) { t.x += x[r].x; t.y += x[r].y; t.z += x[r].z; t.w += x[r].w; }
                    }
                    float4& b = yi ? bsum[1] : bsum[0];
                    b.x += t.x; b.y += t.y; b.z += t.z; b.w += t.w;
                }
                convert(x, nv, hl);
                // every Y' tile of the row tile is released together (after the last MMA of the previous row tile)
                if (pw == 0 && yi == 0) LTTL(3, it, 0);
                umma::mbar_wait(y_empty, (it & 1) ^ 1);
                store(hl, y_img + (size_t)yi * 2 * TILE_BYTES);
            }
            umma::mbar_arrive(y_full);
            if (pw == 0) LTTL(3, it, 1);
#pragma unroll 1
            for (int xi = 0; xi < nxt; ++xi, ++sc) {
                const int s = sc % xstages;
                load_x(x, xi, rfirst, nv);
                if (xi < a.nx) {
                    if (a.x_act[xi] == ACT_SWISH) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            x[r].x = swish_tc<FAST>(x[r].x); x[r].y = swish_tc<FAST>(x[r].y);
                            x[r].z = swish_tc<FAST>(x[r].z); x[r].w = swish_tc<FAST>(x[r].w);
                        }
                    } else if (a.x_act[xi] == ACT_RELU) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            x[r].x = fmaxf(x[r].x, 0.f); x[r].y = fmaxf(x[r].y, 0.f);
                            x[r].z = fmaxf(x[r].z, 0.f); x[r].w = fmaxf(x[r].w, 0.f);
                        }
                    }
                }
                convert(x, nv, hl);
                umma::mbar_wait(&x_empty[s], ((sc / xstages) & 1) ^ 1);
                store(hl, x_stage(s));
                umma::mbar_arrive(&x_full[s]);
                if (pw == 0) LTTL(3, it, xi == 0 ? 2 : 3);
            }
        }
        if (a.bias_partial) {
            for (int yi = 0; yi < a.ny; ++yi)
                *reinterpret_cast<float4*>(a.bias_partial + (((int64_t)blockIdx.x * LT_PROD_WARPS + pw) * a.ny + yi) * 128 + lane * 4) = yi ? bsum[1] : bsum[0];
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == LT_MMA_WARP) umma::tmem_dealloc(tmem, 512);
}

// One launch reduces every accumulator of a wgrad launch over the CTA partials:
//   outs[j].dw[n*lddw + k] (+)= sum_p partial[p][j][n][k]   (n < n_valid, k < k_valid),   db[y][n] (+)= sum_p bias_partial[p][y][n].
// Block = 32 outputs x 8 partial groups (group g sums the partials p = g, g+8, ...; fixed order: deterministic).
struct WgradReduceArgs {
    const float* partial; int nacc; int n_parts;
    WgradTcOut outs[4];
    const float* bias_partial; int ny; int n_bias_parts;
    float* db[2]; int db_accumulate;
};
__global__ void __launch_bounds__(256)
wgrad_tc_reduce_kernel(const WgradReduceArgs a) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int j = blockIdx.y;
    const float* src;
    int64_t stride;
    int n_parts;
    float* dst = nullptr;
    int accumulate;
    const int idx = blockIdx.x * 32 + lane;       // output index inside the 128 x 128 accumulator (or the 128 bias entries)
    if (j < a.nacc) {
        const WgradTcOut& o = a.outs[j];
        const int n = idx >> 7, k = idx & 127;
        src = a.partial + (int64_t)j * 128 * 128 + idx;
        stride = (int64_t)a.nacc * 128 * 128;
        n_parts = a.n_parts;
        accumulate = o.accumulate;
        if (o.dw && n < o.n_valid && k < o.k_valid) dst = o.dw + (int64_t)n * o.lddw + k;
    } else {
        const int y = j - a.nacc;
        src = a.bias_partial + (int64_t)y * 128 + idx;
        stride = (int64_t)a.ny * 128;
        n_parts = a.n_bias_parts;
        accumulate = a.db_accumulate;
        if (blockIdx.x < 4 && a.db[y]) dst = a.db[y] + idx;
    }
    // whole warps share one row n (32 consecutive k), so the test is warp-uniform up to the k_valid edge
    float s = 0.f;
    if (dst) {
        int p = g;
        for (; p + 24 < n_parts; p += 32) {
            const float v0 = src[(int64_t)p * stride], v1 = src[(int64_t)(p + 8) * stride];
            const float v2 = src[(int64_t)(p + 16) * stride], v3 = src[(int64_t)(p + 24) * stride];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; p < n_parts; p += 8) s += src[(int64_t)p * stride];
    }
    red[g][lane] = s;
    __syncthreads();
    if (g == 0 && dst) {
        float t = red[0][lane];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += red[q][lane];
        *dst = accumulate ? *dst + t : t;
    }
}

static int wgrad_tc_grid(int64_t rows) {
    const int64_t tiles = ceil_div<int64_t>(rows > 0 ? rows : 1, 128);
    return (int)(tiles < sm_count() ? tiles : sm_count());
}
size_t wgrad_tc_workspace(int64_t rows) {
    return align_up((size_t)wgrad_tc_grid(rows) * 4 * 128 * 128 * sizeof(float)) + align_up((size_t)wgrad_tc_grid(rows) * LT_PROD_WARPS * 2 * 128 * sizeof(float)) + 256;
}

int launch_wgrad_tc(int precision, WgradTcArgs a, const WgradTcOut* outs, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(a.rows < ((int64_t)1 << 31), "wgrad_tc: row count out of range");
    const int nxt = a.nx + (a.tail ? 1 : 0);
    MGB_REQUIRE(a.ny >= 1 && a.ny <= 2 && a.nx >= 0 && a.nx <= 2 && nxt >= 1 && a.ny * nxt <= 4, "wgrad_tc: at most four accumulators");
    MGB_REQUIRE(a.kt >= 0 && a.kt <= LT_TAIL, "wgrad_tc: tail width must be <= %d", LT_TAIL);
    if (a.rows <= 0) return MGB_OK;
    const int grid = wgrad_tc_grid(a.rows);
    const int nacc = a.ny * nxt;
    Workspace ws(ws_ptr, ws_bytes);
    a.partial = ws.take<float>((size_t)grid * nacc * 128 * 128);
    a.bias_partial = (a.db[0] || a.db[1]) ? ws.take<float>((size_t)grid * LT_PROD_WARPS * a.ny * 128) : nullptr;
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
    WgradReduceArgs r{};
    r.partial = a.partial; r.nacc = nacc; r.n_parts = grid;
    for (int j = 0; j < nacc; ++j) r.outs[j] = outs[j];
    r.bias_partial = a.bias_partial; r.ny = a.ny; r.n_bias_parts = grid * LT_PROD_WARPS;
    r.db[0] = a.db[0]; r.db[1] = a.db[1]; r.db_accumulate = a.db_accumulate;
    wgrad_tc_reduce_kernel<<<dim3(128 * 128 / 32, nacc + (a.bias_partial ? a.ny : 0)), 256, 0, s>>>(r);
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
