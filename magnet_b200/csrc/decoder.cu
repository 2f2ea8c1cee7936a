// magnet_b200 — temporal-bundling decoder + Euler update of MP-PDE (models/mpnn_2d.py:138-162,196-200, models/mpnn.py:139-162,
// 196-200) in one launch per direction:
//     diff = Conv1d(8 -> 1, k2)( act( Conv1d(1 -> 8, k1, stride s1)( h[:, None, :] ) ) )          h [N,128] -> diff [N,tw]
//     out  = u[:, -1:] + cumsum(dt) * diff                                                        dts[j] = (j + 1) dt
// The reference runs it as two cuDNN convolutions over N "images" of 1 x 128, a Swish, a repeat/transpose and two
// element-wise kernels, with [N,8,L1] intermediates in HBM.  Here one warp owns a node: the row sits in shared memory, the
// 8 x L1 hidden values never leave the SM, HBM sees h once in and out [N,tw] once out (HBM-bound: 512 + 4 tw bytes per node).
// Backward recomputes the hidden values, produces dh, d u[:, -1] and per-warp partials of the 217-odd parameter gradients,
// which a second kernel sums in warp order (fixed order, no atomics).
#include "internal.cuh"
#include "dense.cuh"

namespace mgb {

constexpr int DEC_H = 128, DEC_C = 8, DEC_MAXL1 = 60, DEC_MAXK = 16, DEC_WARPS = 8;

__device__ __forceinline__ float dec_act(int act, float z) { return act == ACT_SWISH ? z / (1.0f + expf(-z)) : z; }
__device__ __forceinline__ float dec_act_grad(int act, float z) {
    if (act != ACT_SWISH) return 1.0f;
    const float s = 1.0f / (1.0f + expf(-z));
    return s * fmaf(z, 1.0f - s, 1.0f);
}

struct alignas(16) DecSmem {
    float w1[DEC_C][DEC_MAXK];
    float w2[DEC_C][DEC_MAXK];
    float b1[DEC_C];
    float b2, dt;
    alignas(16) float h[DEC_WARPS][DEC_H];     // float4 row copies
    float z[DEC_WARPS][DEC_C * DEC_MAXL1];        // pre-activations of conv1 (backward: d z)
    float m[DEC_WARPS][DEC_C * DEC_MAXL1];        // activations
    float g[DEC_WARPS][64];                       // backward: d diff
};

__device__ __forceinline__ void dec_load_params(const DecArgs& a, DecSmem& S) {
    for (int i = threadIdx.x; i < DEC_C * DEC_MAXK; i += blockDim.x) {
        const int c = i / DEC_MAXK, k = i % DEC_MAXK;
        S.w1[c][k] = k < a.k1 ? a.w1[c * a.k1 + k] : 0.f;
        S.w2[c][k] = k < a.k2 ? a.w2[c * a.k2 + k] : 0.f;
    }
    if (threadIdx.x < DEC_C) S.b1[threadIdx.x] = a.b1[threadIdx.x];
    if (threadIdx.x == 0) { S.b2 = a.b2[0]; S.dt = a.dt[0]; }
    __syncthreads();
}

// conv1 + activation of one node into the warp's shared rows
__device__ __forceinline__ void dec_hidden(const DecArgs& a, DecSmem& S, int w, int lane, int64_t node) {
    reinterpret_cast<float4*>(S.h[w])[lane] = reinterpret_cast<const float4*>(a.h + node * DEC_H)[lane];
    __syncwarp();
    for (int idx = lane; idx < DEC_C * a.L1; idx += 32) {
        const int c = idx / a.L1, p = idx - c * a.L1;
        float z = S.b1[c];
        const float* hp = S.h[w] + p * a.s1;
#pragma unroll 4
        for (int k = 0; k < a.k1; ++k) z = fmaf(S.w1[c][k], hp[k], z);
        S.z[w][idx] = z;
        S.m[w][idx] = dec_act(a.act, z);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(DEC_WARPS * 32) decoder_fwd_kernel(const DecArgs a, float* __restrict__ out) {
    __shared__ DecSmem S;
    dec_load_params(a, S);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t node = (int64_t)blockIdx.x * DEC_WARPS + w; node < a.n; node += (int64_t)gridDim.x * DEC_WARPS) {
        dec_hidden(a, S, w, lane, node);
        const float ul = a.u[node * a.ldu + a.u_col];
        for (int j = lane; j < a.tw; j += 32) {
            float d = S.b2;
            for (int c = 0; c < DEC_C; ++c) {
                const float* mp = S.m[w] + c * a.L1 + j;
#pragma unroll 2
                for (int k = 0; k < a.k2; ++k) d = fmaf(S.w2[c][k], mp[k], d);
            }
            out[node * a.tw + j] = fmaf((float)(j + 1) * S.dt, d, ul);
        }
        __syncwarp();
    }
}

// per-warp partial layout: [dW1 8*16 | db1 8 | dW2 8*16 | db2 1] (padded kernels: entries k >= k1 / k2 stay zero)
constexpr int DEC_NPARAM = DEC_C * DEC_MAXK + DEC_C + DEC_C * DEC_MAXK + 1;

__global__ void __launch_bounds__(DEC_WARPS * 32) decoder_bwd_kernel(const DecArgs a, const float* __restrict__ dout, float* __restrict__ dh,
                                                                     float* __restrict__ du, int lddu, float* __restrict__ partial) {
    __shared__ DecSmem S;
    dec_load_params(a, S);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float aw1[4] = {0.f, 0.f, 0.f, 0.f}, aw2[4] = {0.f, 0.f, 0.f, 0.f}, ab1 = 0.f, ab2 = 0.f;      // (c, k) = index lane + 32 i
    for (int64_t node = (int64_t)blockIdx.x * DEC_WARPS + w; node < a.n; node += (int64_t)gridDim.x * DEC_WARPS) {
        dec_hidden(a, S, w, lane, node);
        // d diff[j] = dout[j] * dts[j];  d u_last = sum_j dout[j]
        float su = 0.f;
        for (int j = lane; j < a.tw; j += 32) {
            const float d = dout[node * a.tw + j];
            su += d;
            S.g[w][j] = d * (float)(j + 1) * S.dt;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) su += __shfl_xor_sync(0xffffffffu, su, o);
        if (lane == 0 && du) du[node * lddu + a.u_col] = su;
        __syncwarp();
        // conv2 backward: weight gradients from the activations, then d z = (d m) act'(z) in place
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = lane + 32 * i, c = idx / DEC_MAXK, k = idx % DEC_MAXK;
            if (k < a.k2) {
                const float* mp = S.m[w] + c * a.L1 + k;
                float s = 0.f;
                for (int j = 0; j < a.tw; ++j) s = fmaf(S.g[w][j], mp[j], s);
                aw2[i] += s;
            }
        }
        if (lane == 0) {
            float s = 0.f;
            for (int j = 0; j < a.tw; ++j) s += S.g[w][j];
            ab2 += s;
        }
        for (int idx = lane; idx < DEC_C * a.L1; idx += 32) {
            const int c = idx / a.L1, p = idx - c * a.L1;
            const int j0 = p - a.k2 + 1 > 0 ? p - a.k2 + 1 : 0, j1 = p < a.tw - 1 ? p : a.tw - 1;
            float s = 0.f;
            for (int j = j0; j <= j1; ++j) s = fmaf(S.g[w][j], S.w2[c][p - j], s);
            S.z[w][idx] = s * dec_act_grad(a.act, S.z[w][idx]);
        }
        __syncwarp();
        // conv1 backward: weight / bias gradients, d h
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = lane + 32 * i, c = idx / DEC_MAXK, k = idx % DEC_MAXK;
            if (k < a.k1) {
                const float* zp = S.z[w] + c * a.L1;
                float s = 0.f;
                for (int p = 0; p < a.L1; ++p) s = fmaf(zp[p], S.h[w][p * a.s1 + k], s);
                aw1[i] += s;
            }
        }
        if (lane < DEC_C) {
            const float* zp = S.z[w] + lane * a.L1;
            float s = 0.f;
            for (int p = 0; p < a.L1; ++p) s += zp[p];
            ab1 += s;
        }
        float4 o;
        float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = lane * 4 + e;
            int p0 = (i - a.k1 + a.s1) / a.s1;               // smallest p with p*s1 + k1 - 1 >= i
            if (i - a.k1 + 1 <= 0) p0 = 0;
            int p1 = i / a.s1;
            if (p1 > a.L1 - 1) p1 = a.L1 - 1;
            float s = 0.f;
            for (int p = p0; p <= p1; ++p) {
                const int k = i - p * a.s1;
#pragma unroll
                for (int c = 0; c < DEC_C; ++c) s = fmaf(S.z[w][c * a.L1 + p], S.w1[c][k], s);
            }
            ov[e] = s;
        }
        reinterpret_cast<float4*>(dh + node * DEC_H)[lane] = o;
        __syncwarp();
    }
    float* pw = partial + ((size_t)blockIdx.x * DEC_WARPS + w) * DEC_NPARAM;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pw[lane + 32 * i] = aw1[i];
        pw[DEC_C * DEC_MAXK + DEC_C + lane + 32 * i] = aw2[i];
    }
    if (lane < DEC_C) pw[DEC_C * DEC_MAXK + lane] = ab1;
    if (lane == 0) pw[DEC_NPARAM - 1] = ab2;
}

__global__ void decoder_reduce_kernel(const float* __restrict__ partial, int n_warps, int k1, int k2, float* __restrict__ dw1,
                                      float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= DEC_NPARAM) return;
    float s = 0.f;
    for (int wv = 0; wv < n_warps; ++wv) s += partial[(size_t)wv * DEC_NPARAM + i];
    if (i < DEC_C * DEC_MAXK) {
        const int c = i / DEC_MAXK, k = i % DEC_MAXK;
        if (k < k1) dw1[c * k1 + k] = s;
    } else if (i < DEC_C * DEC_MAXK + DEC_C) {
        db1[i - DEC_C * DEC_MAXK] = s;
    } else if (i < DEC_NPARAM - 1) {
        const int j = i - DEC_C * DEC_MAXK - DEC_C, c = j / DEC_MAXK, k = j % DEC_MAXK;
        if (k < k2) dw2[c * k2 + k] = s;
    } else {
        db2[0] = s;
    }
}

static int dec_grid(int64_t n) {
    const int64_t blocks = ceil_div<int64_t>(n, DEC_WARPS);
    const int64_t cap = (int64_t)sm_count() * 4;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

static int dec_check(DecArgs& a, int hidden) {
    MGB_REQUIRE(hidden == DEC_H, "decoder: built for hidden_features = 128 (got %d)", hidden);
    MGB_REQUIRE(a.k1 >= 1 && a.k1 <= DEC_MAXK && a.k2 >= 1 && a.k2 <= DEC_MAXK && a.s1 >= 1, "decoder: kernel sizes 1..16 (got %d, %d)", a.k1, a.k2);
    a.L1 = (DEC_H - a.k1) / a.s1 + 1;
    MGB_REQUIRE(a.L1 <= DEC_MAXL1 && a.L1 - a.k2 + 1 == a.tw && a.tw <= 64, "decoder: conv shapes give %d outputs, time_window is %d", a.L1 - a.k2 + 1, a.tw);
    MGB_REQUIRE(((uintptr_t)a.h % 16) == 0, "decoder: h must be 16-byte aligned");
    return MGB_OK;
}

size_t decoder_bwd_workspace(int64_t n) { return align_up((size_t)dec_grid(n) * DEC_WARPS * DEC_NPARAM * sizeof(float)) + 256; }

int decoder_fwd(DecArgs a, int hidden, float* out, cudaStream_t s) {
    MGB_TRY(dec_check(a, hidden));
    if (a.n <= 0) return MGB_OK;
    decoder_fwd_kernel<<<dec_grid(a.n), DEC_WARPS * 32, 0, s>>>(a, out);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

int decoder_bwd(DecArgs a, int hidden, const float* dout, float* dh, float* du, int lddu, float* dw1, float* db1, float* dw2, float* db2,
                void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    MGB_TRY(dec_check(a, hidden));
    MGB_REQUIRE(((uintptr_t)dh % 16) == 0, "decoder: dh must be 16-byte aligned");
    const int grid = dec_grid(a.n);
    Workspace ws(ws_ptr, ws_bytes);
    float* partial = ws.take<float>((size_t)grid * DEC_WARPS * DEC_NPARAM);
    MGB_WS_CHECK(ws);
    decoder_bwd_kernel<<<grid, DEC_WARPS * 32, 0, s>>>(a, dout, dh, du, lddu, partial);
    MGB_LAUNCH_CHECK();
    decoder_reduce_kernel<<<ceil_div(DEC_NPARAM, 128), 128, 0, s>>>(partial, grid * DEC_WARPS, a.k1, a.k2, dw1, db1, dw2, db2);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
