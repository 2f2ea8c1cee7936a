// magnet_b200 — helpers shared by the tcgen05 kernels (gnn_edge_tc.cu, linear_tc.cu).
#pragma once
#include "umma.cuh"
#include "dense.cuh"

namespace mgb {

constexpr int TILE_BYTES = 128 * 256;   // one [128][128] bf16 operand image (two 128-byte-swizzle blocks)

// ---- activations ------------------------------------------------------------------------------------
// FAST (bf16 contract): sigmoid through one MUFU.TANH; precise: ex2.approx + rcp.approx (2 MUFU, ~2 ulp).
template <bool FAST>
__device__ __forceinline__ float sigmoid_tc(float z) {
    if (FAST) {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
        return fmaf(0.5f, t, 0.5f);
    }
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
template <bool FAST>
__device__ __forceinline__ float swish_tc(float z) {
    if (FAST) {
        float t;
        const float h = 0.5f * z;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        return fmaf(h, t, h);
    }
    return z * sigmoid_tc<false>(z);
}
template <bool FAST>
__device__ __forceinline__ float swish_grad_tc(float z) {
    const float s = sigmoid_tc<FAST>(z);
    return s * fmaf(z, 1.0f - s, 1.0f);
}

// two floats -> bf16x2 (hi) and the bf16x2 of the residuals (lo); ~3 instructions per element
__device__ __forceinline__ void split2_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = umma::pack_bf16(a, b);
    lo = umma::pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

// the same split into two fp16 values: 22 significant bits (|x - hi - lo| <= 2^-22 |x| while lo stays a normal number,
// i.e. for |x| >~ 0.03; an absolute 3e-8 below that); |x| must stay below 65504
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void load_w2_image(unsigned char* w_img, const unsigned char* src, uint32_t bytes, uint64_t* wbar) {
    // one bulk async copy (TMA engine; no tensor map is needed for a pre-swizzled image)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(wbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     umma::smem_u32(w_img)),
                 "l"(src), "r"(bytes), "r"(umma::smem_u32(wbar))
                 : "memory");
}


// generic activation on the tensor-core paths: 0 none, 1 ReLU, 2 Swish
template <bool FAST>
__device__ __forceinline__ float act_tc(int act, float z) {
    if (act == ACT_RELU) return fmaxf(z, 0.f);
    if (act == ACT_SWISH) return swish_tc<FAST>(z);
    return z;
}
template <bool FAST>
__device__ __forceinline__ float act_grad_tc(int act, float z) {
    if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
    if (act == ACT_SWISH) return swish_grad_tc<FAST>(z);
    return 1.f;
}

}  // namespace mgb
