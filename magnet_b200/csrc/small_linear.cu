// magnet_b200 — Linear layers with a handful of input features and 128 outputs: the first Linear of MAgNet's Encoder MLPs
// (models/magnet_gnn.py:20-35: 13 node / 12 edge features -> 128) and of MP-PDE's embedding (models/mpnn_2d.py:130-131).
// A K <= 16 contraction has no use for a 128x128x16 GEMM tile pipeline: the work is writing (forward) or reading (backward)
// 512 bytes per row.  One warp owns a row at a time, a lane owns four output channels (one 16-byte access per row), the
// weight slice of a lane lives in registers:
//   forward   y[r, :] = act(x[r, :] Wt + b)            HBM: 4 K + 512 (+ 512 for the saved pre-activation) bytes per row
//   backward  g = dy . act'(y_pre);  dx[r, k] = sum_n g[n] W[n, k] (warp reduction);  dW[n, k] += g[n] x[r, k];  db[n] += g[n]
//             dW / db accumulate in registers over the rows of a warp and are summed over the warps in warp order by a second
//             kernel (fixed order, no atomics).
#include "internal.cuh"
#include "dense.cuh"

namespace mgb {

constexpr int SL_MAXK = 16, SL_WARPS = 8;

template <int K>
__global__ void __launch_bounds__(SL_WARPS * 32) small_linear_fwd_kernel(const float* __restrict__ x, int64_t rows, int ldx, const float* __restrict__ wt /*[K][128]*/,
                                                                        const float* __restrict__ bias, int act, float* __restrict__ y,
                                                                        float* __restrict__ y_pre) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float4 wk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = reinterpret_cast<const float4*>(wt + k * 128)[lane];
    const float4 b = bias ? reinterpret_cast<const float4*>(bias)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = (int64_t)blockIdx.x * SL_WARPS + w; r < rows; r += (int64_t)gridDim.x * SL_WARPS) {
        const float xv = lane < K ? x[r * ldx + lane] : 0.f;
        float4 acc = b;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float xk = __shfl_sync(0xffffffffu, xv, k);
            acc.x = fmaf(xk, wk[k].x, acc.x); acc.y = fmaf(xk, wk[k].y, acc.y);
            acc.z = fmaf(xk, wk[k].z, acc.z); acc.w = fmaf(xk, wk[k].w, acc.w);
        }
        if (y_pre) reinterpret_cast<float4*>(y_pre + r * 128)[lane] = acc;
        acc.x = act_apply(act, acc.x); acc.y = act_apply(act, acc.y); acc.z = act_apply(act, acc.z); acc.w = act_apply(act, acc.w);
        reinterpret_cast<float4*>(y + r * 128)[lane] = acc;
    }
}

// partial layout per warp: [K][128] weight gradients (k-major, like wt) followed by [128] bias gradients.
// DX = false (the input is data: raw edge / node features) drops the weight slice from the registers and keeps the gradient
// rows of FOUR rows per warp in flight: with one row (1 KB) per warp and ~120 registers the kernel had 16 KB in flight per SM
// against the ~43 KB that 6.4 TB/s x 1 us needs, and ran at 0.4 of the HBM bound on the 12.6 M-row edge encoder.
template <int K, bool DX>
__global__ void __launch_bounds__(SL_WARPS * 32) small_linear_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y_pre, int act,
                                                                        const float* __restrict__ x, int64_t rows, int ldx,
                                                                        const float* __restrict__ wt, float* __restrict__ dx,
                                                                        float* __restrict__ partial) {
    constexpr int U = DX ? 1 : 4;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float4 wk[DX ? K : 1], gw[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (DX) wk[k] = reinterpret_cast<const float4*>(wt + k * 128)[lane];
        gw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 gb = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t stride = (int64_t)gridDim.x * SL_WARPS;
    for (int64_t r0 = (int64_t)blockIdx.x * SL_WARPS + w; r0 < rows; r0 += stride * U) {
        float4 gq[U], zq[U];
        float xq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t r = r0 + u * stride;
            const bool ok = r < rows;
            gq[u] = ok ? reinterpret_cast<const float4*>(dy + r * 128)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
            zq[u] = (ok && act != ACT_NONE) ? reinterpret_cast<const float4*>(y_pre + r * 128)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
            xq[u] = (ok && lane < K) ? x[r * ldx + lane] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t r = r0 + u * stride;
            float4 g = gq[u];
            if (act != ACT_NONE) {
                const float4 z = zq[u];
                g.x *= act_grad(act, z.x); g.y *= act_grad(act, z.y); g.z *= act_grad(act, z.z); g.w *= act_grad(act, z.w);
            }
            const float xv = xq[u];
            gb.x += g.x; gb.y += g.y; gb.z += g.z; gb.w += g.w;
            float mine = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float xk = __shfl_sync(0xffffffffu, xv, k);
                gw[k].x = fmaf(g.x, xk, gw[k].x); gw[k].y = fmaf(g.y, xk, gw[k].y);
                gw[k].z = fmaf(g.z, xk, gw[k].z); gw[k].w = fmaf(g.w, xk, gw[k].w);
                if (DX) {
                    float s = (g.x * wk[k].x + g.y * wk[k].y) + (g.z * wk[k].z + g.w * wk[k].w);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == k) mine = s;
                }
            }
            if (DX && lane < K && r < rows) dx[r * ldx + lane] = mine;
        }
    }
    float* p = partial + ((size_t)blockIdx.x * SL_WARPS + w) * (K + 1) * 128;
#pragma unroll
    for (int k = 0; k < K; ++k) reinterpret_cast<float4*>(p + k * 128)[lane] = gw[k];
    reinterpret_cast<float4*>(p + K * 128)[lane] = gb;
}

// dw[n][k] (+)= sum over warps of partial[w][k][n];  db[n] (+)= sum of partial[w][K][n]
__global__ void small_linear_reduce_kernel(const float* __restrict__ partial, int n_warps, int K, float* __restrict__ dw, int lddw,
                                           float* __restrict__ db, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (K + 1) * 128) return;
    float s = 0.f;
    for (int wv = 0; wv < n_warps; ++wv) s += partial[(size_t)wv * (K + 1) * 128 + i];
    const int k = i / 128, n = i % 128;
    if (k < K) {
        if (dw) dw[n * lddw + k] = accumulate ? dw[n * lddw + k] + s : s;
    } else if (db) {
        db[n] = accumulate ? db[n] + s : s;
    }
}

static int sl_grid(int64_t rows) {
    const int64_t blocks = ceil_div<int64_t>(rows, SL_WARPS);
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

bool small_linear_ok(int in_features, int out_features) { return in_features >= 1 && in_features <= SL_MAXK && out_features == 128; }

#define SL_DISPATCH(K, CALL)                                                                           \
    switch (K) {                                                                                       \
        case 1: { constexpr int KK = 1; CALL; } break;   case 2: { constexpr int KK = 2; CALL; } break;    \
        case 3: { constexpr int KK = 3; CALL; } break;   case 4: { constexpr int KK = 4; CALL; } break;    \
        case 5: { constexpr int KK = 5; CALL; } break;   case 6: { constexpr int KK = 6; CALL; } break;    \
        case 7: { constexpr int KK = 7; CALL; } break;   case 8: { constexpr int KK = 8; CALL; } break;    \
        case 9: { constexpr int KK = 9; CALL; } break;   case 10: { constexpr int KK = 10; CALL; } break;  \
        case 11: { constexpr int KK = 11; CALL; } break; case 12: { constexpr int KK = 12; CALL; } break;  \
        case 13: { constexpr int KK = 13; CALL; } break; case 14: { constexpr int KK = 14; CALL; } break;  \
        case 15: { constexpr int KK = 15; CALL; } break; default: { constexpr int KK = 16; CALL; } break;  \
    }

int small_linear_fwd(const float* x, int64_t rows, int K, const float* wt, const float* bias, int act, float* y, float* y_pre, cudaStream_t s) {
    if (rows <= 0) return MGB_OK;
    MGB_REQUIRE(((uintptr_t)wt % 16) == 0 && ((uintptr_t)y % 16) == 0 && (!bias || ((uintptr_t)bias % 16) == 0) && (!y_pre || ((uintptr_t)y_pre % 16) == 0),
                "small_linear_fwd: wt / bias / y must be 16-byte aligned");
    const int grid = sl_grid(rows);
    SL_DISPATCH(K, (small_linear_fwd_kernel<KK><<<grid, SL_WARPS * 32, 0, s>>>(x, rows, K, wt, bias, act, y, y_pre)));
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

size_t small_linear_bwd_workspace(int64_t rows, int K) {
    return align_up((size_t)sl_grid(rows) * SL_WARPS * (K + 1) * 128 * sizeof(float)) + align_up((size_t)K * 128 * sizeof(float)) + 256;
}

// w: [128][K] (PyTorch layout); transposed into the workspace for the kernels
int small_linear_bwd(const float* dy, const float* y_pre, int act, const float* x, int64_t rows, int K, const float* w, float* dx, float* dw,
                     float* db, int accumulate, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(((uintptr_t)dy % 16) == 0 && (act == ACT_NONE || ((uintptr_t)y_pre % 16) == 0), "small_linear_bwd: dy / y_pre must be 16-byte aligned");
    const int grid = sl_grid(rows);
    Workspace ws(ws_ptr, ws_bytes);
    float* partial = ws.take<float>((size_t)grid * SL_WARPS * (K + 1) * 128);
    float* wt = ws.take<float>((size_t)K * 128);
    MGB_WS_CHECK(ws);
    MGB_TRY(launch_transpose(w, 128, K, K, wt, 128, s));
    if (dx) { SL_DISPATCH(K, (small_linear_bwd_kernel<KK, true><<<grid, SL_WARPS * 32, 0, s>>>(dy, y_pre, act, x, rows, K, wt, dx, partial))); }
    else { SL_DISPATCH(K, (small_linear_bwd_kernel<KK, false><<<grid, SL_WARPS * 32, 0, s>>>(dy, y_pre, act, x, rows, K, wt, dx, partial))); }
    MGB_LAUNCH_CHECK();
    small_linear_reduce_kernel<<<ceil_div((K + 1) * 128, 256), 256, 0, s>>>(partial, grid * SL_WARPS, K, dw, K, db, accumulate);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
