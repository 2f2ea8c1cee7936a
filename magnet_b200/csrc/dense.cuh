// magnet_b200 — row-wise dense building blocks (fp32 FFMA path): C = act(A·B + bias) (+R),
// the matching weight-gradient reduction, column sums, and activation helpers.
//
// These serve the node-level stages of both layer flavours (update nets, encoders, decoders,
// projector) and are the arithmetic fall-back that locks the fp32 1e-5 contract.  The
// per-edge contractions use the fused kernels in gnn_layer.cu / interaction.cu.
#pragma once
#include "common.cuh"

namespace mgb {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SWISH = 2 };

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float act_apply(int act, float x) {
    if (act == ACT_RELU) return fmaxf(x, 0.0f);
    if (act == ACT_SWISH) return x * sigmoid_f(x);
    return x;
}
// derivative of the activation with respect to its pre-activation input z
__device__ __forceinline__ float act_grad(int act, float z) {
    if (act == ACT_RELU) return z > 0.0f ? 1.0f : 0.0f;
    if (act == ACT_SWISH) { float s = sigmoid_f(z); return s * (1.0f + z * (1.0f - s)); }
    return 1.0f;
}

// Logical A operand: up to four row-major segments concatenated along K; optional element-wise
// prologue  a := a * act'(pre)  (used by the data-gradient GEMMs) or a := act(a) (recompute).
struct ASpec {
    const float* p[4];
    int ld[4];
    int k[4];
    int nseg;
    const float* pre;   // same shape/ld as segment 0 (only with nseg == 1)
    int pre_act;        // a *= act_grad(pre_act, pre)
    int self_act;       // a = act_apply(self_act, a)   (activation recomputed from a saved pre-activation)
};

struct GemmArgs {
    ASpec a;
    const float* b;      // [K, N] row-major
    int ldb;
    const float* bias;   // [N] or null
    const float* residual;
    int ldr;
    float* c;            // activated output [M, N]
    int ldc;
    float* c_pre;        // optional pre-activation copy
    int ldcp;
    int act;
    int M, N, K;
    int accumulate;      // c += result (act must be ACT_NONE)
};

int launch_gemm(const GemmArgs& g, cudaStream_t s);

// dW[n][k] = sum_r Y'[r][n] * A'[r][k],   Y' = dy * act'(y_pre) (optional),  A' per ASpec.
struct WgradArgs {
    const float* dy; int lddy;
    const float* y_pre; int y_act;      // optional prologue on dy (same ld as dy)
    ASpec a;
    int rows, N, K;                     // N = out features, K = in features (sum of segments)
    float* dw; int lddw;                // [N, K] row-major (PyTorch Linear.weight layout)
    float* db;                          // [N] or null
    int accumulate;                     // dw/db += (parameter shared between several calls)
};
size_t wgrad_workspace_bytes(int rows, int N, int K);
int launch_wgrad(const WgradArgs& w, void* ws, size_t ws_bytes, cudaStream_t s);

// out[c] (+)= sum_r x[r][c] * (pre ? act'(pre[r][c]) : 1)
size_t colsum_workspace_bytes(int rows, int cols);
int launch_colsum(const float* x, int ld, const float* pre, int act, int rows, int cols, float* out, int accumulate,
                  void* ws, size_t ws_bytes, cudaStream_t s);

int launch_transpose(const float* in, int rows, int cols, int ld_in, float* out, int ld_out, cudaStream_t s);

// y = LayerNorm(x) * gamma + beta, eps 1e-5, biased variance (nn.LayerNorm(128))
int launch_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* rstd_mean /*[rows][2]*/,
                         int64_t rows, int cols, cudaStream_t s, const float* residual = nullptr);
size_t layernorm_bwd_workspace_bytes(int64_t rows, int cols);
int launch_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* rstd_mean, float* dx,
                         float* dgamma, float* dbeta, int accumulate_params, int64_t rows, int cols, void* ws,
                         size_t ws_bytes, cudaStream_t s);

// ------------------------------------------------------------------------------------------
// shared-memory tile contraction used by the fused edge kernels (FFMA path):
//   acc[i][j] += sum_k A[(te*4+i)][k] * B[k][col(j)],  64 x 128 output tile, 256 threads,
//   A: [64][LDA] row-major in smem, B: [KDIM][128] row-major in smem.
// thread (te = tid/16, tn = tid%16) owns rows te*4..+3 and columns tn*4..+3, 64+tn*4..+3.
// ------------------------------------------------------------------------------------------
template <int KDIM, int LDA>
__device__ __forceinline__ void tile_fma_64x128(const float* __restrict__ As, const float* __restrict__ Bs, float (&acc)[4][8]) {
    const int te = threadIdx.x >> 4, tn = threadIdx.x & 15;
    const float* a0 = As + (te * 4) * LDA;
    const float* b0 = Bs + tn * 4;
#pragma unroll 2
    for (int k0 = 0; k0 < KDIM; k0 += 4) {
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * LDA + k0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 w0 = *reinterpret_cast<const float4*>(b0 + (k0 + kk) * 128);
            const float4 w1 = *reinterpret_cast<const float4*>(b0 + (k0 + kk) * 128 + 64);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                acc[i][0] = fmaf(av, w0.x, acc[i][0]);
                acc[i][1] = fmaf(av, w0.y, acc[i][1]);
                acc[i][2] = fmaf(av, w0.z, acc[i][2]);
                acc[i][3] = fmaf(av, w0.w, acc[i][3]);
                acc[i][4] = fmaf(av, w1.x, acc[i][4]);
                acc[i][5] = fmaf(av, w1.y, acc[i][5]);
                acc[i][6] = fmaf(av, w1.z, acc[i][6]);
                acc[i][7] = fmaf(av, w1.w, acc[i][7]);
            }
        }
    }
}

}  // namespace mgb
