// magnet_b200 — flat-buffer Adam (one kernel for the ~150 parameter tensors of a model).
//
// Reference: configure_optimizers, models/magnet_gnn.py:378-386 and models/mpnn_2d.py:205-213 —
// torch.optim.Adam(lr, weight_decay) (L2 penalty added to the gradient, not decoupled) under a StepLR schedule.
// Same update, same order of operations as torch.optim.Adam's single-tensor path:
//   g   = grad * grad_scale + weight_decay * p            (grad_scale: 1/world after a sum all-reduce)
//   m   = beta1 m + (1 - beta1) g
//   v   = beta2 v + (1 - beta2) g^2
//   p  -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// The learning rate is a plain argument: the StepLR decay is host arithmetic.
#include "internal.cuh"

namespace mgb {

__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 float step_size, float beta1, float beta2, float eps, float weight_decay, float inv_sqrt_bc2, float grad_scale) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* pe = &pp.x; const float* ge = &gg.x; float* me = &mm.x; float* ve = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = fmaf(weight_decay, pe[k], ge[k] * grad_scale);
            me[k] = fmaf(beta1, me[k], (1.0f - beta1) * gr);
            ve[k] = fmaf(beta2, ve[k], (1.0f - beta2) * gr * gr);
            pe[k] -= step_size * me[k] / (sqrtf(ve[k]) * inv_sqrt_bc2 + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail (n not a multiple of 4)
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gr = fmaf(weight_decay, p[i], g[i] * grad_scale);
        const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gr);
        const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gr * gr);
        m[i] = mi; v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

int adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
              double weight_decay, int64_t step, double grad_scale, cudaStream_t s) {
    MGB_REQUIRE(n >= 0 && step >= 1, "adam_step: n >= 0 and step >= 1 (got %lld, %lld)", (long long)n, (long long)step);
    MGB_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, "adam_step: buffers must be 16-byte aligned");
    if (n == 0) return MGB_OK;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const int64_t want = ceil_div<int64_t>(n / 4 + 1, 256);
    const int64_t cap = (int64_t)sm_count() * 8;
    adam_step_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(p, g, m, v, n, (float)(lr / bc1), (float)beta1, (float)beta2, (float)eps,
                                                                         (float)weight_decay, (float)(1.0 / sqrt(bc2)), (float)grad_scale);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
