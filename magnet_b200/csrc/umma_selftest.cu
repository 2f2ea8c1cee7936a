// magnet_b200 — device self-test of the tcgen05 building blocks in umma.cuh (test hook, not a product path).
// One CTA computes D[m][n] = sum_k A(m,k) B(n,k) for a single 128x128x128 bf16 tile with fp32 accumulation in
// TMEM, for every combination of K-major / MN-major operand descriptors over the SAME swizzled tile image.
#include "internal.cuh"
#include "umma.cuh"

namespace mgb {

struct UmmaTestArgs {
    const float* a;   // [128][128] storage rows x cols: a_mn == 0 -> a[m][k], a_mn == 1 -> a[k][m]
    const float* b;   // b_mn == 0 -> b[n][k], b_mn == 1 -> b[k][n]
    float* d;         // [128][128]  d[m][n]
    int a_mn, b_mn;
    int lbo_mn, sbo_mn;   // MN-major descriptor overrides (0 = defaults of umma.cuh)
};

__global__ void __launch_bounds__(128) umma_selftest_kernel(const UmmaTestArgs t) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const uint32_t raw = umma::smem_u32(smem_raw);
    const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
    unsigned char* tile_a = smem_raw + pad;
    unsigned char* tile_b = tile_a + 128 * 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int idx = tid; idx < 128 * 128; idx += 128) {
        const int r = idx >> 7, c = idx & 127;
        *reinterpret_cast<__nv_bfloat16*>(tile_a + umma::tile_off(128, r, c)) = __float2bfloat16_rn(t.a[idx]);
        *reinterpret_cast<__nv_bfloat16*>(tile_b + umma::tile_off(128, r, c)) = __float2bfloat16_rn(t.b[idx]);
    }
    umma::fence_async_smem();
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::fence_barrier_init();
    }
    if (warp == 0) umma::tmem_alloc(&tmem_base, 256);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (t.a_mn == 2) {
        // A in tensor memory (TS form): thread m writes row m of A as 64 packed bf16 pairs into columns [128, 192)
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
            float w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j)
                w[j] = __uint_as_float(umma::pack_bf16(t.a[tid * 128 + 2 * (c0 + j)], t.a[tid * 128 + 2 * (c0 + j) + 1]));
            umma::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 128u + (uint32_t)c0, w);
        }
        umma::tc_fence_before();
        __syncthreads();
        umma::tc_fence_after();
    }
    if (tid == 0 && t.a_mn == 2) {
        const uint32_t idesc = umma::idesc_bf16(128, 128, 0, t.b_mn);
        const uint32_t sb = umma::smem_u32(tile_b);
        const uint32_t lbo = t.lbo_mn ? (uint32_t)t.lbo_mn : 128u * 128u, sbo = t.sbo_mn ? (uint32_t)t.sbo_mn : 1024u;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            const uint64_t db = t.b_mn ? umma::desc_sw128(sb + k * 2048, lbo, sbo) : umma::desc_kmajor(sb, k);
            umma::mma_bf16_ts(tmem, tmem + 128u + (uint32_t)(k * 8), db, idesc, k > 0);
        }
        umma::mma_commit(&bar);
    } else if (tid == 0) {
        const uint32_t idesc = umma::idesc_bf16(128, 128, t.a_mn, t.b_mn);
        const uint32_t sa = umma::smem_u32(tile_a), sb = umma::smem_u32(tile_b);
        const uint32_t lbo = t.lbo_mn ? (uint32_t)t.lbo_mn : 128u * 128u, sbo = t.sbo_mn ? (uint32_t)t.sbo_mn : 1024u;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            const uint64_t da = t.a_mn ? umma::desc_sw128(sa + k * 2048, lbo, sbo) : umma::desc_kmajor(sa, k);
            const uint64_t db = t.b_mn ? umma::desc_sw128(sb + k * 2048, lbo, sbo) : umma::desc_kmajor(sb, k);
            umma::mma_bf16(tmem, da, db, idesc, k > 0);
        }
        umma::mma_commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) t.d[(warp * 32 + lane) * 128 + c0 + i] = v[i];
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

int umma_selftest(const float* a, const float* b, int a_mn, int b_mn, int lbo_mn, int sbo_mn, float* d, cudaStream_t s) {
    UmmaTestArgs t{a, b, d, a_mn, b_mn, lbo_mn, sbo_mn};
    const size_t smem = 2 * 128 * 256 + 1024;
    MGB_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, s>>>(t);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
