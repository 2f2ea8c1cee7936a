// magnet_b200 — MP-PDE message-passing layer (GNN_Layer), forward and backward, FFMA path.
//
// Reference: GNN_Layer.forward/message/update — models/mpnn_2d.py:65-90 (models/mpnn.py:65-90),
// mean aggregation aggr='mean' (:46) and PyG InstanceNorm (:63,70).
//
//   m_e  = Swish(W2 Swish(W1 [x_i, x_j, u_i-u_j, pos_i-pos_j, var_i] + b1) + b2)      i = dst, j = src
//   agg  = mean_{e -> i} m_e                                                         (0 if none)
//   out  = x + Swish(W4 Swish(W3 [x, agg, var] + b3) + b4);   y = InstanceNorm(out, batch)
//
// B200 design:
//  * the first Linear is factorised per node: W1·[...] = P[i] + Q[j] with
//        P = [x,u,pos,var]·Wa + b1,   Q = [x,u,pos]·Wb      (one node-level GEMM, 256 outputs)
//    so nothing of width 269 is ever gathered or materialised per edge; the per-edge work is
//    Swish(P[dst]+Q[src]) followed by ONE 128x128 contraction — the only dense per-edge
//    contraction left — and a segmented mean over the destination-sorted CSR.
//  * edges are processed in fixed tiles of 64 positions of the dst-sorted order; complete
//    segments are reduced inside the tile (no atomics), segments cut by a tile boundary go
//    through head/tail partial rows and a deterministic fix-up pass.
//  * backward recomputes the messages from P, Q (nothing of size [E, .] is saved by forward),
//    reduces dP by destination in-kernel and dQ by source through the transposed CSR.
#include "internal.cuh"

namespace mgb {

constexpr int H = 128;          // hidden width of the reference configs (hidden_features / latent_dim)
constexpr int TE = 64;          // edge positions per tile
constexpr int LDH = H + 4;      // smem row pitch of the edge tiles (conflict-free broadcast reads)

// ------------------------------------------------------------------------------------------
// weight packing: W1 [H, 2H+tw+dp+nv] -> Wcat_t [Kc, 2H], Wcat [2H, Kc], bcat [2H];  Kc = H+tw+dp+nv
// ------------------------------------------------------------------------------------------
__global__ void pack_w1_kernel(const float* __restrict__ W1, const float* __restrict__ b1, int tw, int dp, int nv,
                               float* __restrict__ wcat_t, float* __restrict__ wcat, float* __restrict__ bcat) {
    const int Kc = H + tw + dp + nv, K1 = 2 * H + tw + dp + nv;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2 * H) bcat[idx] = idx < H ? b1[idx] : 0.f;
    if (idx >= Kc * 2 * H) return;
    const int k = idx / (2 * H), c = idx - k * 2 * H;   // Xc feature k, output column c (P: c<H, Q: c>=H)
    const int n = c < H ? c : c - H;
    float v;
    if (k < H) v = W1[n * K1 + (c < H ? k : H + k)];
    else if (k < H + tw + dp) { float w = W1[n * K1 + 2 * H + (k - H)]; v = c < H ? w : -w; }
    else v = c < H ? W1[n * K1 + 2 * H + (k - H)] : 0.f;
    wcat_t[k * 2 * H + c] = v;
    wcat[c * Kc + k] = v;
}

// dWcat [2H, Kc] -> dW1 [H, K1]
__global__ void unpack_dw1_kernel(const float* __restrict__ dwcat, int tw, int dp, int nv, float* __restrict__ dW1,
                                  int accumulate) {
    const int Kc = H + tw + dp + nv, K1 = 2 * H + tw + dp + nv;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * K1) return;
    const int n = idx / K1, col = idx - n * K1;
    float v;
    if (col < H) v = dwcat[n * Kc + col];
    else if (col < 2 * H) v = dwcat[(H + n) * Kc + (col - H)];
    else if (col < 2 * H + tw + dp) v = dwcat[n * Kc + (col - H)] - dwcat[(H + n) * Kc + (col - H)];
    else v = dwcat[n * Kc + (col - H)];
    dW1[idx] = accumulate ? dW1[idx] + v : v;
}

// ------------------------------------------------------------------------------------------
// segmented reduction of a [TE][LDH] smem tile over the dst-sorted positions (no atomics)
// thread c (< H) owns column c.  MEAN: divide complete segments by their length.
// ------------------------------------------------------------------------------------------
template <bool MEAN>
__device__ __forceinline__ void tile_segment_reduce(const float* __restrict__ tile, const int* __restrict__ dst_s, int ne,
                                                    int64_t e0, const int32_t* __restrict__ rowptr, float* __restrict__ out,
                                                    int ld_out, float* __restrict__ part_head, float* __restrict__ part_tail,
                                                    int64_t tile_id, int c) {
    int cur = dst_s[0];
    float acc = 0.f;
    const int64_t e1 = e0 + ne;
    for (int e = 0; e <= ne; ++e) {
        const int d = e < ne ? dst_s[e] : -1;
        if (d != cur) {
            const int64_t s0 = rowptr[cur], s1 = rowptr[cur + 1];
            if (s0 >= e0 && s1 <= e1) out[(int64_t)cur * ld_out + c] = MEAN ? acc / (float)(s1 - s0) : acc;
            else if (s0 < e0) part_head[tile_id * H + c] = acc;
            else part_tail[tile_id * H + c] = acc;
            cur = d;
            acc = 0.f;
        }
        if (e < ne) acc += tile[e * LDH + c];
    }
}

// one block (H threads) per tile boundary: finishes the segment that ENDS in tile t but began earlier
template <bool MEAN>
__global__ void __launch_bounds__(H)
segment_fixup_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ dstv, int64_t n_edges,
                     const float* __restrict__ part_head, const float* __restrict__ part_tail, float* __restrict__ out, int ld_out) {
    const int64_t t = (int64_t)blockIdx.x + 1;
    const int64_t e0 = t * TE;
    if (e0 >= n_edges) return;
    const int node = dstv[e0];
    const int64_t s0 = rowptr[node], s1 = rowptr[node + 1];
    if (!(s0 < e0 && s1 <= e0 + TE)) return;     // not open at the start, or continues into a later tile
    const int c = threadIdx.x;
    const int64_t t0 = s0 / TE;
    float acc = part_tail[t0 * H + c];
    for (int64_t tt = t0 + 1; tt <= t; ++tt) acc += part_head[tt * H + c];
    out[(int64_t)node * ld_out + c] = MEAN ? acc / (float)(s1 - s0) : acc;
}

// ------------------------------------------------------------------------------------------
// fused edge kernel, forward:  gather P[dst], Q[src] -> Swish -> x W2^T + b2 -> Swish -> segmented mean
// ------------------------------------------------------------------------------------------
struct EdgeFwdArgs {
    const float* pq;       // [N][2H]: P | Q
    const int32_t* rowptr; // [N+1]   dst-sorted CSR
    const int32_t* dstv;   // [E]
    const int32_t* srcv;   // [E]
    int64_t n_edges;
    const float* w2t;      // [H][H]  (k-major: w2t[k][n] = W2[n][k])
    const float* b2;       // [H]
    float* agg;            // [N][H]  (pre-zeroed: isolated nodes stay 0)
    float* part_head;      // [tiles][H]
    float* part_tail;      // [tiles][H]
};

constexpr size_t EDGE_FWD_SMEM = (size_t)(H * H + TE * LDH) * sizeof(float) + 2 * TE * sizeof(int);

__global__ void __launch_bounds__(256, 2) gnn_edge_fwd_kernel(const EdgeFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ws = reinterpret_cast<float*>(smem_raw);
    float* Hs = Ws + H * H;
    int* dst_s = reinterpret_cast<int*>(Hs + TE * LDH);
    int* src_s = dst_s + TE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int te = tid >> 4, tn = tid & 15;
    for (int i = tid; i < H * H / 4; i += 256)
        reinterpret_cast<float4*>(Ws)[i] = reinterpret_cast<const float4*>(a.w2t)[i];
    const float4 bias0 = *reinterpret_cast<const float4*>(a.b2 + tn * 4);
    const float4 bias1 = *reinterpret_cast<const float4*>(a.b2 + 64 + tn * 4);
    const int64_t n_tiles = ceil_div<int64_t>(a.n_edges, TE);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * TE;
        const int ne = (int)((a.n_edges - e0) < (int64_t)TE ? (a.n_edges - e0) : (int64_t)TE);
        __syncthreads();                                  // previous tile fully consumed
        if (tid < TE) {
            dst_s[tid] = tid < ne ? a.dstv[e0 + tid] : -1;
            src_s[tid] = tid < ne ? a.srcv[e0 + tid] : -1;
        }
        __syncthreads();
        // gather: warp w builds rows w, w+8, ...; lane covers 4 consecutive channels
#pragma unroll
        for (int i = 0; i < TE / 8; ++i) {
            const int e = warp + 8 * i;
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < ne) {
                const float4 p = *reinterpret_cast<const float4*>(a.pq + (int64_t)dst_s[e] * (2 * H) + lane * 4);
                const float4 q = *reinterpret_cast<const float4*>(a.pq + (int64_t)src_s[e] * (2 * H) + H + lane * 4);
                h.x = act_apply(ACT_SWISH, p.x + q.x); h.y = act_apply(ACT_SWISH, p.y + q.y);
                h.z = act_apply(ACT_SWISH, p.z + q.z); h.w = act_apply(ACT_SWISH, p.w + q.w);
            }
            *reinterpret_cast<float4*>(Hs + e * LDH + lane * 4) = h;
        }
        __syncthreads();
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        tile_fma_64x128<H, LDH>(Hs, Ws, acc);
        __syncthreads();                                  // all reads of Hs done
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 o0, o1;
            o0.x = act_apply(ACT_SWISH, acc[i][0] + bias0.x); o0.y = act_apply(ACT_SWISH, acc[i][1] + bias0.y);
            o0.z = act_apply(ACT_SWISH, acc[i][2] + bias0.z); o0.w = act_apply(ACT_SWISH, acc[i][3] + bias0.w);
            o1.x = act_apply(ACT_SWISH, acc[i][4] + bias1.x); o1.y = act_apply(ACT_SWISH, acc[i][5] + bias1.y);
            o1.z = act_apply(ACT_SWISH, acc[i][6] + bias1.z); o1.w = act_apply(ACT_SWISH, acc[i][7] + bias1.w);
            *reinterpret_cast<float4*>(Hs + (te * 4 + i) * LDH + tn * 4) = o0;
            *reinterpret_cast<float4*>(Hs + (te * 4 + i) * LDH + 64 + tn * 4) = o1;
        }
        __syncthreads();
        if (tid < H) tile_segment_reduce<true>(Hs, dst_s, ne, e0, a.rowptr, a.agg, H, a.part_head, a.part_tail, tile, tid);
    }
}

// ------------------------------------------------------------------------------------------
// fused edge kernel, backward (recompute): per tile
//   h1 = Swish(z1), z1 = P[dst]+Q[src];  z2 = h1 W2^T + b2;  dz2 = (dagg[dst]/deg) * Swish'(z2)
//   dW2 += dz2^T h1;  db2 += colsum(dz2);  dh1 = dz2 W2;  dz1 = dh1 * Swish'(z1)
//   dz1 -> global [E][H] (for the by-source reduction) and segmented SUM by destination -> dP
// ------------------------------------------------------------------------------------------
struct EdgeBwdArgs {
    const float* pq;
    const int32_t* rowptr;
    const int32_t* dstv;
    const int32_t* srcv;
    int64_t n_edges;
    const float* w2t;      // [H][H] k-major
    const float* w2;       // [H][H] PyTorch layout [out][in]
    const float* b2;
    const float* dagg;     // [N][ld_dagg]
    int ld_dagg;
    float* dz1;            // [E][H]
    float* dpq;            // [N][2H]; this kernel writes columns [0,H) (dP), pre-zeroed
    float* part_head;
    float* part_tail;
    float* dw2_partial;    // [grid][H][H]
    float* db2_partial;    // [grid][H]
};

constexpr size_t EDGE_BWD_SMEM = (size_t)(2 * H * H + 2 * TE * LDH) * sizeof(float) + 2 * TE * sizeof(int);

__global__ void __launch_bounds__(256, 1) gnn_edge_bwd_kernel(const EdgeBwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Wts = reinterpret_cast<float*>(smem_raw);    // W2^T  (forward recompute)
    float* Wos = Wts + H * H;                           // W2    (data gradient)
    float* H1s = Wos + H * H;                           // h1, later dz1
    float* DZs = H1s + TE * LDH;                        // dz2
    int* dst_s = reinterpret_cast<int*>(DZs + TE * LDH);
    int* src_s = dst_s + TE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int te = tid >> 4, tn = tid & 15;
    for (int i = tid; i < H * H / 4; i += 256) {
        reinterpret_cast<float4*>(Wts)[i] = reinterpret_cast<const float4*>(a.w2t)[i];
        reinterpret_cast<float4*>(Wos)[i] = reinterpret_cast<const float4*>(a.w2)[i];
    }
    const float4 bias0 = *reinterpret_cast<const float4*>(a.b2 + tn * 4);
    const float4 bias1 = *reinterpret_cast<const float4*>(a.b2 + 64 + tn * 4);
    float dw[8][8];       // dW2[n][k] patch: n in {te*4.., 64+te*4..}, k in {tn*4.., 64+tn*4..}
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dw[i][j] = 0.f;
    float db = 0.f;       // threads < H: column sum of dz2
    const int64_t n_tiles = ceil_div<int64_t>(a.n_edges, TE);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * TE;
        const int ne = (int)((a.n_edges - e0) < (int64_t)TE ? (a.n_edges - e0) : (int64_t)TE);
        __syncthreads();
        if (tid < TE) {
            dst_s[tid] = tid < ne ? a.dstv[e0 + tid] : -1;
            src_s[tid] = tid < ne ? a.srcv[e0 + tid] : -1;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < TE / 8; ++i) {
            const int e = warp + 8 * i;
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < ne) {
                const float4 p = *reinterpret_cast<const float4*>(a.pq + (int64_t)dst_s[e] * (2 * H) + lane * 4);
                const float4 q = *reinterpret_cast<const float4*>(a.pq + (int64_t)src_s[e] * (2 * H) + H + lane * 4);
                h.x = act_apply(ACT_SWISH, p.x + q.x); h.y = act_apply(ACT_SWISH, p.y + q.y);
                h.z = act_apply(ACT_SWISH, p.z + q.z); h.w = act_apply(ACT_SWISH, p.w + q.w);
            }
            *reinterpret_cast<float4*>(H1s + e * LDH + lane * 4) = h;
        }
        __syncthreads();
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        tile_fma_64x128<H, LDH>(H1s, Wts, acc);           // z2 (without bias)
        // dz2 = dagg[dst] / deg * Swish'(z2)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = te * 4 + i;
            float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
            if (e < ne) {
                const int d = dst_s[e];
                const float inv = 1.0f / (float)(a.rowptr[d + 1] - a.rowptr[d]);
                const float4 g0 = *reinterpret_cast<const float4*>(a.dagg + (int64_t)d * a.ld_dagg + tn * 4);
                const float4 g1 = *reinterpret_cast<const float4*>(a.dagg + (int64_t)d * a.ld_dagg + 64 + tn * 4);
                o0.x = g0.x * inv * act_grad(ACT_SWISH, acc[i][0] + bias0.x); o0.y = g0.y * inv * act_grad(ACT_SWISH, acc[i][1] + bias0.y);
                o0.z = g0.z * inv * act_grad(ACT_SWISH, acc[i][2] + bias0.z); o0.w = g0.w * inv * act_grad(ACT_SWISH, acc[i][3] + bias0.w);
                o1.x = g1.x * inv * act_grad(ACT_SWISH, acc[i][4] + bias1.x); o1.y = g1.y * inv * act_grad(ACT_SWISH, acc[i][5] + bias1.y);
                o1.z = g1.z * inv * act_grad(ACT_SWISH, acc[i][6] + bias1.z); o1.w = g1.w * inv * act_grad(ACT_SWISH, acc[i][7] + bias1.w);
            }
            *reinterpret_cast<float4*>(DZs + e * LDH + tn * 4) = o0;
            *reinterpret_cast<float4*>(DZs + e * LDH + 64 + tn * 4) = o1;
        }
        __syncthreads();
        // dW2 patch += dz2^T h1 ; db2 += colsum(dz2)
#pragma unroll 4
        for (int e = 0; e < TE; ++e) {
            const float4 y0 = *reinterpret_cast<const float4*>(DZs + e * LDH + te * 4);
            const float4 y1 = *reinterpret_cast<const float4*>(DZs + e * LDH + 64 + te * 4);
            const float4 h0 = *reinterpret_cast<const float4*>(H1s + e * LDH + tn * 4);
            const float4 h1 = *reinterpret_cast<const float4*>(H1s + e * LDH + 64 + tn * 4);
            const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
            const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dw[i][j] = fmaf(yv[i], hv[j], dw[i][j]);
        }
        if (tid < H) {
            float s = 0.f;
#pragma unroll 8
            for (int e = 0; e < TE; ++e) s += DZs[e * LDH + tid];
            db += s;
        }
        // dh1 = dz2 W2
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        tile_fma_64x128<H, LDH>(DZs, Wos, acc);
        __syncthreads();                                  // h1 no longer needed
        // dz1 = dh1 * Swish'(z1); z1 re-gathered (L1/L2 resident)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = te * 4 + i;
            float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
            if (e < ne) {
                const float* pr = a.pq + (int64_t)dst_s[e] * (2 * H);
                const float* qr = a.pq + (int64_t)src_s[e] * (2 * H) + H;
                const float4 p0 = *reinterpret_cast<const float4*>(pr + tn * 4), p1 = *reinterpret_cast<const float4*>(pr + 64 + tn * 4);
                const float4 q0 = *reinterpret_cast<const float4*>(qr + tn * 4), q1 = *reinterpret_cast<const float4*>(qr + 64 + tn * 4);
                o0.x = acc[i][0] * act_grad(ACT_SWISH, p0.x + q0.x); o0.y = acc[i][1] * act_grad(ACT_SWISH, p0.y + q0.y);
                o0.z = acc[i][2] * act_grad(ACT_SWISH, p0.z + q0.z); o0.w = acc[i][3] * act_grad(ACT_SWISH, p0.w + q0.w);
                o1.x = acc[i][4] * act_grad(ACT_SWISH, p1.x + q1.x); o1.y = acc[i][5] * act_grad(ACT_SWISH, p1.y + q1.y);
                o1.z = acc[i][6] * act_grad(ACT_SWISH, p1.z + q1.z); o1.w = acc[i][7] * act_grad(ACT_SWISH, p1.w + q1.w);
                float* g = a.dz1 + (e0 + e) * H;
                *reinterpret_cast<float4*>(g + tn * 4) = o0;
                *reinterpret_cast<float4*>(g + 64 + tn * 4) = o1;
            }
            *reinterpret_cast<float4*>(H1s + e * LDH + tn * 4) = o0;
            *reinterpret_cast<float4*>(H1s + e * LDH + 64 + tn * 4) = o1;
        }
        __syncthreads();
        if (tid < H) tile_segment_reduce<false>(H1s, dst_s, ne, e0, a.rowptr, a.dpq, 2 * H, a.part_head, a.part_tail, tile, tid);
    }
    // per-CTA partial of dW2 / db2 (reduced in fixed order by the host-side second stage)
    float* out = a.dw2_partial + (int64_t)blockIdx.x * H * H;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = i < 4 ? te * 4 + i : 64 + te * 4 + (i - 4);
        *reinterpret_cast<float4*>(out + n * H + tn * 4) = make_float4(dw[i][0], dw[i][1], dw[i][2], dw[i][3]);
        *reinterpret_cast<float4*>(out + n * H + 64 + tn * 4) = make_float4(dw[i][4], dw[i][5], dw[i][6], dw[i][7]);
    }
    if (tid < H) a.db2_partial[(int64_t)blockIdx.x * H + tid] = db;
}

__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ partial, int n_parts, int64_t count, float* __restrict__ out, int accumulate) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int p = 0; p < n_parts; ++p) s += partial[(int64_t)p * count + i];
    out[i] = accumulate ? out[i] + s : s;
}

// out[j][0:H) = sum over q in [rowptr_t[j], rowptr_t[j+1]) of rows[pos_t[q]][0:H)   (one warp per node)
__global__ void __launch_bounds__(256)
gather_sum_rows_kernel(const float* __restrict__ rows, const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ pos_t,
                       int64_t n_nodes, float* __restrict__ out, int ld_out) {
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_nodes) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int s = rowptr_t[j], e = rowptr_t[j + 1];
    for (int q = s; q < e; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(rows + (int64_t)pos_t[q] * H + lane * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + j * ld_out + lane * 4) = acc;
}

// out[i][0:H) = 0 for every node without in-edges (rowptr[i] == rowptr[i+1]).  The tensor-core edge kernels store the rows of
// all other nodes in full, so this replaces a memset of the whole [N,128] tensor (15-25 us per layer at 131 k nodes).
__global__ void __launch_bounds__(256)
zero_isolated_rows_kernel(const int32_t* __restrict__ rowptr, int64_t n_nodes, float* __restrict__ out, int ld_out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool iso = i < n_nodes && rowptr[i] == rowptr[i + 1];
    uint32_t m = __ballot_sync(0xffffffffu, iso);
    const int64_t base = i - lane;
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        *reinterpret_cast<float4*>(out + (base + b) * ld_out + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------
// InstanceNorm over the nodes of each graph (graphs are contiguous row ranges given by gptr)
// ------------------------------------------------------------------------------------------
constexpr int IN_ROWS = 128;   // rows per block

// mode 0: sum x          mode 1: sum (x-mean)^2        mode 2: sums of dy and dy*yhat (two outputs)
template <int MODE>
__global__ void __launch_bounds__(H)
inorm_partial_kernel(const float* __restrict__ x, const float* __restrict__ aux, const int64_t* __restrict__ gptr,
                     const float* __restrict__ mean, float* __restrict__ partial, int chunks) {
    const int g = blockIdx.y, j = blockIdx.x, c = threadIdx.x;
    const int64_t r0 = gptr[g] + (int64_t)j * IN_ROWS, r1 = min(gptr[g + 1], r0 + IN_ROWS);
    float s0 = 0.f, s1 = 0.f;
    const float m = MODE == 1 ? mean[g * H + c] : 0.f;
    for (int64_t r = r0; r < r1; ++r) {
        const float v = x[r * H + c];
        if (MODE == 0) s0 += v;
        if (MODE == 1) { const float dd = v - m; s0 += dd * dd; }
        if (MODE == 2) { s0 += v; s1 += v * aux[r * H + c]; }
    }
    const int64_t o = ((int64_t)g * chunks + j) * H + c;
    partial[o] = s0;
    if (MODE == 2) partial[(int64_t)gridDim.y * chunks * H + o] = s1;
}

// mode 0 -> mean; mode 1 -> rstd = 1/sqrt(var+eps); mode 2 -> two means
template <int MODE>
__global__ void __launch_bounds__(H)
inorm_finalize_kernel(const float* __restrict__ partial, const int64_t* __restrict__ gptr, int chunks, int n_graphs,
                      float* __restrict__ out0, float* __restrict__ out1) {
    const int g = blockIdx.x, c = threadIdx.x;
    const int64_t cnt = gptr[g + 1] - gptr[g];
    const int used = (int)ceil_div<int64_t>(cnt, IN_ROWS);
    const float norm = (float)(cnt > 1 ? cnt : 1);
    float s0 = 0.f, s1 = 0.f;
    for (int j = 0; j < used; ++j) {
        const int64_t o = ((int64_t)g * chunks + j) * H + c;
        s0 += partial[o];
        if (MODE == 2) s1 += partial[(int64_t)n_graphs * chunks * H + o];
    }
    if (MODE == 0) out0[g * H + c] = s0 / norm;
    if (MODE == 1) out0[g * H + c] = 1.0f / sqrtf(s0 / norm + 1e-5f);
    if (MODE == 2) { out0[g * H + c] = s0 / norm; out1[g * H + c] = s1 / norm; }
}

// fwd: y = (x - mean) * rstd      bwd: dx = rstd * (dy - m1 - yhat * m2)
template <bool BWD>
__global__ void __launch_bounds__(H)
inorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ yhat, const int64_t* __restrict__ gptr,
                   const float* __restrict__ m0, const float* __restrict__ m1, const float* __restrict__ rstd, float* __restrict__ out) {
    const int g = blockIdx.y, j = blockIdx.x, c = threadIdx.x;
    const int64_t r0 = gptr[g] + (int64_t)j * IN_ROWS, r1 = min(gptr[g + 1], r0 + IN_ROWS);
    const float a = m0[g * H + c], rs = rstd[g * H + c];
    const float b = BWD ? m1[g * H + c] : 0.f;
    for (int64_t r = r0; r < r1; ++r) {
        if (BWD) out[r * H + c] = rs * (x[r * H + c] - a - yhat[r * H + c] * b);
        else out[r * H + c] = (x[r * H + c] - a) * rs;
    }
}

size_t inorm_workspace_bytes(int n_graphs, int max_nodes) {
    int chunks = ceil_div(max_nodes > 0 ? max_nodes : 1, IN_ROWS);
    return align_up((size_t)2 * n_graphs * chunks * H * sizeof(float)) + 2 * align_up((size_t)n_graphs * H * sizeof(float)) + 512;
}

// y = InstanceNorm(x); rstd_out [G][H] is kept for the backward pass
int instance_norm_fwd(const float* x, const int64_t* gptr, int n_graphs, int max_nodes, float* y, float* rstd_out, void* ws_ptr,
                      size_t ws_bytes, cudaStream_t s) {
    if (n_graphs == 0) return MGB_OK;
    const int chunks = ceil_div(max_nodes > 0 ? max_nodes : 1, IN_ROWS);
    Workspace ws(ws_ptr, ws_bytes);
    float* partial = ws.take<float>((size_t)2 * n_graphs * chunks * H);
    float* mean = ws.take<float>((size_t)n_graphs * H);
    MGB_WS_CHECK(ws);
    dim3 grid(chunks, n_graphs);
    inorm_partial_kernel<0><<<grid, H, 0, s>>>(x, nullptr, gptr, nullptr, partial, chunks);
    MGB_LAUNCH_CHECK();
    inorm_finalize_kernel<0><<<n_graphs, H, 0, s>>>(partial, gptr, chunks, n_graphs, mean, nullptr);
    MGB_LAUNCH_CHECK();
    inorm_partial_kernel<1><<<grid, H, 0, s>>>(x, nullptr, gptr, mean, partial, chunks);
    MGB_LAUNCH_CHECK();
    inorm_finalize_kernel<1><<<n_graphs, H, 0, s>>>(partial, gptr, chunks, n_graphs, rstd_out, nullptr);
    MGB_LAUNCH_CHECK();
    inorm_apply_kernel<false><<<grid, H, 0, s>>>(x, nullptr, gptr, mean, nullptr, rstd_out, y);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

int instance_norm_bwd(const float* dy, const float* y, const float* rstd, const int64_t* gptr, int n_graphs, int max_nodes,
                      float* dx, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    if (n_graphs == 0) return MGB_OK;
    const int chunks = ceil_div(max_nodes > 0 ? max_nodes : 1, IN_ROWS);
    Workspace ws(ws_ptr, ws_bytes);
    float* partial = ws.take<float>((size_t)2 * n_graphs * chunks * H);
    float* m1 = ws.take<float>((size_t)n_graphs * H);
    float* m2 = ws.take<float>((size_t)n_graphs * H);
    MGB_WS_CHECK(ws);
    dim3 grid(chunks, n_graphs);
    inorm_partial_kernel<2><<<grid, H, 0, s>>>(dy, y, gptr, nullptr, partial, chunks);
    MGB_LAUNCH_CHECK();
    inorm_finalize_kernel<2><<<n_graphs, H, 0, s>>>(partial, gptr, chunks, n_graphs, m1, m2);
    MGB_LAUNCH_CHECK();
    inorm_apply_kernel<true><<<grid, H, 0, s>>>(dy, y, gptr, m1, m2, rstd, dx);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
// packed weight block (floats): wcat_t [Kc][2H] | wcat [2H][Kc] | bcat [2H] | w2t [H][H] | w3t [K3][H] | w4t [H][H]
struct GnnPacked {
    float *wcat_t, *wcat, *bcat, *w2t, *w3t, *w4t;
    float* w2img;   // two swizzled bf16 [128][128] images of W2 (hi | lo) for the tcgen05 edge kernels
    float* img_pq;  // weight tiles (hi|lo images, H*H floats each) of the node-level Linears for linear_tc.cu:
    float* img_w3;  //   Wcat rows P / Q x columns [0,128);  W3 columns [0,128) / [128,256);  W4
    float* img_w4;
    float* wtm;     // W3[:, 0:128], W3[:, 128:256], W4 in tensor-memory order (bf16 hi | lo, 16384 words each) for node_update_tc.cu
    float* wtm_pq;  // linear_ts.cu, same order: Wcat rows P / Q x columns [0,128)   (P | Q = [x,..] Wcat^T)
    float* wtm_dc;  //   W3^T output blocks [0,128) / [128,256)                        (dc = d1' W3)
    float* wtm_dx;  //   Wcat^T K-tiles P / Q                                          (dx = dP Wcat[:128] + dQ Wcat[128:])
    float* wtm_d1;  //   W4^T                                                          (d1 = d0' W4)
};
static size_t packed_layout(const GnnLayerShape& sh, float* base, GnnPacked* p) {
    size_t off = 0;
    auto take = [&](size_t n) { float* r = base ? base + off : nullptr; off += (n + 63) / 64 * 64; return r; };
    float* a = take((size_t)sh.Kc() * 2 * H);
    float* b = take((size_t)2 * H * sh.Kc());
    float* c = take(2 * H);
    float* d = take((size_t)H * H);
    float* e = take((size_t)sh.K3() * H);
    float* f = take((size_t)H * H);
    float* g = take((size_t)H * H);      // 2 images x 128 x 128 bf16 = H*H floats
    float* i1 = take((size_t)2 * H * H);
    float* i2 = take((size_t)2 * H * H);
    float* i3 = take((size_t)H * H);
    float* i4 = take((size_t)3 * H * H);
    float* i5 = take((size_t)2 * H * H);
    float* i6 = take((size_t)2 * H * H);
    float* i7 = take((size_t)2 * H * H);
    float* i8 = take((size_t)H * H);
    if (p) { p->wtm = i4; p->wtm_pq = i5; p->wtm_dc = i6; p->wtm_dx = i7; p->wtm_d1 = i8; }
    if (p) { p->wcat_t = a; p->wcat = b; p->bcat = c; p->w2t = d; p->w3t = e; p->w4t = f; p->w2img = g;
             p->img_pq = i1; p->img_w3 = i2; p->img_w4 = i3; }
    return off;
}

size_t gnn_layer_packed_floats(int tw, int dp, int nv) {
    GnnLayerShape sh{0, 0, tw, dp, nv, 0, 0, 0};
    return packed_layout(sh, nullptr, nullptr);
}

int gnn_layer_pack(const float* W1, const float* b1, const float* W2, const float* W3, const float* W4, int tw, int dp,
                   int nv, float* packed, cudaStream_t s) {
    GnnLayerShape sh{0, 0, tw, dp, nv, 0, 0, 0};
    GnnPacked p;
    packed_layout(sh, packed, &p);
    pack_w1_kernel<<<ceil_div(sh.Kc() * 2 * H, 256), 256, 0, s>>>(W1, b1, tw, dp, nv, p.wcat_t, p.wcat, p.bcat);
    MGB_LAUNCH_CHECK();
    MGB_TRY(launch_transpose(W2, H, H, H, p.w2t, H, s));
    MGB_TRY(launch_transpose(W3, H, sh.K3(), sh.K3(), p.w3t, H, s));
    MGB_TRY(launch_transpose(W4, H, H, H, p.w4t, H, s));
    MGB_TRY(pack_w2_image(W2, p.w2img, s));
    MGB_TRY(pack_weight_tile(p.wcat, sh.Kc(), 2 * H, sh.Kc(), 0, 0, p.img_pq, s));
    MGB_TRY(pack_weight_tile(p.wcat, sh.Kc(), 2 * H, sh.Kc(), H, 0, p.img_pq + H * H, s));
    MGB_TRY(pack_weight_tile(W3, sh.K3(), H, sh.K3(), 0, 0, p.img_w3, s));
    MGB_TRY(pack_weight_tile(W3, sh.K3(), H, sh.K3(), 0, H, p.img_w3 + H * H, s));
    MGB_TRY(pack_weight_tile(W4, H, H, H, 0, 0, p.img_w4, s));
    MGB_TRY(pack_weight_tmem_bf16(W3, sh.K3(), 0, p.wtm, s));
    MGB_TRY(pack_weight_tmem_bf16(W3, sh.K3(), H, p.wtm + H * H, s));
    MGB_TRY(pack_weight_tmem_bf16(W4, H, 0, p.wtm + 2 * H * H, s));
    MGB_TRY(pack_weight_tmem_bf16(p.wcat, sh.Kc(), 0, p.wtm_pq, s));                              // A[n][k] = Wcat[n][k], P rows
    MGB_TRY(pack_weight_tmem_bf16(p.wcat + (size_t)H * sh.Kc(), sh.Kc(), 0, p.wtm_pq + H * H, s));   // Q rows
    MGB_TRY(pack_weight_tmem_bf16(p.w3t, H, 0, p.wtm_dc, s));                                     // A[j][n] = W3[n][j], j < 128
    MGB_TRY(pack_weight_tmem_bf16(p.w3t + (size_t)H * H, H, 0, p.wtm_dc + H * H, s));                //   j in [128, 256)
    MGB_TRY(pack_weight_tmem_bf16(p.wcat_t, 2 * H, 0, p.wtm_dx, s));                              // A[k][n] = Wcat[n][k], n < 128
    MGB_TRY(pack_weight_tmem_bf16(p.wcat_t, 2 * H, H, p.wtm_dx + H * H, s));                       //   n in [128, 256)
    MGB_TRY(pack_weight_tmem_bf16(p.w4t, H, 0, p.wtm_d1, s));                                      // A[k][n] = W4[n][k]
    return MGB_OK;
}

// developer switch (MGB_TS_LINEARS=0: the P | Q Linear and the two-tile data gradients through linear_tc.cu, the round-1 path)
static bool ts_linears() {
    static const bool on = [] { const char* e = getenv("MGB_TS_LINEARS"); return !(e && e[0] == '0'); }();
    return on;
}
// developer switch (MGB_NODE_UPDATE=0: update_net_1 / update_net_2 as two launches of linear_tc.cu, the round-1 path)
static bool fused_node_update() {
    static const bool on = [] { const char* e = getenv("MGB_NODE_UPDATE"); return !(e && e[0] == '0'); }();
    return on;
}

static int edge_grid(int64_t n_tiles, int ctas_per_sm) {
    int64_t g = (int64_t)sm_count() * ctas_per_sm;
    return (int)(n_tiles < g ? (n_tiles > 0 ? n_tiles : 1) : g);
}

size_t gnn_layer_fwd_workspace(int64_t n_nodes, int64_t n_edges, int n_graphs, int max_nodes) {
    int64_t tiles = ceil_div<int64_t>(n_edges > 0 ? n_edges : 1, TE);
    return 2 * align_up((size_t)tiles * H * sizeof(float)) + align_up((size_t)n_nodes * H * sizeof(float)) * 2 +
           align_up(edge_fwd_tc_workspace(n_edges)) + inorm_workspace_bytes(n_graphs, max_nodes) + 4096;
}

int gnn_layer_fwd(const GnnLayerShape& sh, const GnnFwdIO& io, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(sh.n_nodes < ((int64_t)1 << 31) && sh.n_edges < ((int64_t)1 << 31), "gnn_layer: sizes out of range");
    const int N = (int)sh.n_nodes;
    GnnPacked p;
    packed_layout(sh, const_cast<float*>(io.packed), &p);
    Workspace ws(ws_ptr, ws_bytes);
    const int64_t tiles = ceil_div<int64_t>(sh.n_edges > 0 ? sh.n_edges : 1, TE);
    float* part_head = ws.take<float>((size_t)tiles * H);
    float* part_tail = ws.take<float>((size_t)tiles * H);
    float* y1 = ws.take<float>((size_t)N * H);
    float* out = ws.take<float>((size_t)N * H);
    const size_t tc_bytes = edge_fwd_tc_workspace(sh.n_edges);
    char* tc_ws = ws.take<char>(tc_bytes);
    MGB_WS_CHECK(ws);
    MGB_REQUIRE(sh.precision >= 0 && sh.precision <= 2, "gnn_layer: unknown precision %d", sh.precision);
    if (N == 0) return MGB_OK;
    const int kt_pq = sh.tw + sh.dp + sh.nv;
    const bool tc_nodes = sh.precision != 0 && kt_pq <= 16 && sh.nv <= 16;
    // 1. P | Q = [x,u,pos,var] Wcat^T + [b1 | 0]
    if (tc_nodes) {
        LinTcArgs a{};
        a.src[0] = io.x; a.ld[0] = H; a.nk = 1;
        a.tsrc[0] = io.u; a.tld[0] = sh.tw; a.tk[0] = sh.tw;
        a.tsrc[1] = io.pos; a.tld[1] = sh.dp; a.tk[1] = sh.dp;
        a.tsrc[2] = io.var; a.tld[2] = sh.nv; a.tk[2] = sh.nv;
        a.kt = kt_pq; a.wtail = p.wcat + H; a.wt_sn = sh.Kc(); a.wt_st = 1;
        a.wimg = p.img_pq; a.nm = 2; a.tile_of[0][0] = 0; a.tile_of[1][0] = 1;
        a.bias = p.bcat; a.act = ACT_NONE; a.y = io.pq; a.ldy = 2 * H; a.rows = N;
        if (ts_linears()) { a.wimg = p.wtm_pq; MGB_TRY(launch_linear_ts(sh.precision, a, s)); }
        else MGB_TRY(launch_linear_tc(sh.precision, a, s));
    } else {
        GemmArgs g{};
        g.a.p[0] = io.x; g.a.ld[0] = H; g.a.k[0] = H;
        g.a.p[1] = io.u; g.a.ld[1] = sh.tw; g.a.k[1] = sh.tw;
        g.a.p[2] = io.pos; g.a.ld[2] = sh.dp; g.a.k[2] = sh.dp;
        g.a.p[3] = io.var; g.a.ld[3] = sh.nv; g.a.k[3] = sh.nv;
        g.a.nseg = 4;
        g.b = p.wcat_t; g.ldb = 2 * H; g.bias = p.bcat;
        g.c = io.pq; g.ldc = 2 * H; g.act = ACT_NONE;
        g.M = N; g.N = 2 * H; g.K = sh.Kc();
        MGB_TRY(launch_gemm(g, s));
    }
    // 2. fused edge kernel + boundary fix-up
    if (sh.n_edges > 0 && sh.precision != 0) {
        zero_isolated_rows_kernel<<<(unsigned)ceil_div<int64_t>(N, 256), 256, 0, s>>>(io.rowptr, N, io.agg, H);
        MGB_LAUNCH_CHECK();
    } else {
        MGB_CUDA(cudaMemsetAsync(io.agg, 0, (size_t)N * H * sizeof(float), s));
    }
    if (sh.n_edges > 0 && sh.precision != 0) {
        MGB_TRY(launch_edge_fwd_tc(sh.precision, io.pq, io.rowptr, io.dstv, io.srcv, sh.n_edges, p.w2img, io.b2, io.agg,
                                   tc_ws, tc_bytes, s));
    } else if (sh.n_edges > 0) {
        EdgeFwdArgs a{io.pq, io.rowptr, io.dstv, io.srcv, sh.n_edges, p.w2t, io.b2, io.agg, part_head, part_tail};
        MGB_CUDA(cudaFuncSetAttribute(gnn_edge_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EDGE_FWD_SMEM));
        {
            ProfScope prof(PROF_EDGE_FWD, s);
            gnn_edge_fwd_kernel<<<edge_grid(tiles, 2), 256, EDGE_FWD_SMEM, s>>>(a);
        }
        MGB_LAUNCH_CHECK();
        if (tiles > 1) {
            segment_fixup_kernel<true><<<(unsigned)(tiles - 1), H, 0, s>>>(io.rowptr, io.dstv, sh.n_edges, part_head, part_tail, io.agg, H);
            MGB_LAUNCH_CHECK();
        }
    }
    // 3. update net
    if (tc_nodes && sh.nv <= 4 && fused_node_update()) {
        // both Linears, the Swish between them and the residual in one launch (node_update_tc.cu)
        NodeUpdateArgs a{};
        a.x = io.x; a.agg = io.agg; a.rows = N; a.var = io.var; a.nv = sh.nv;
        a.wimg = p.wtm; a.w3tail = p.w3t + (size_t)2 * H * H; a.wt_sn = 1; a.wt_st = H;
        a.b3 = io.b3; a.b4 = io.b4; a.y1_pre = io.y1_pre; a.y2_pre = io.y2_pre; a.out = out;
        MGB_TRY(launch_node_update_tc(sh.precision, a, s));
    } else if (tc_nodes) {
        LinTcArgs a{};
        a.src[0] = io.x; a.ld[0] = H; a.src[1] = io.agg; a.ld[1] = H; a.nk = 2;
        a.tsrc[0] = io.var; a.tld[0] = sh.nv; a.tk[0] = sh.nv; a.kt = sh.nv;
        a.wtail = p.w3t + (size_t)2 * H * H; a.wt_sn = 1; a.wt_st = H;
        a.wimg = p.img_w3; a.nm = 1; a.tile_of[0][0] = 0; a.tile_of[0][1] = 1;
        // only the pre-activation is written (it is what the backward pass needs); update_net_2 applies Swish on load
        a.bias = io.b3; a.act = ACT_NONE; a.y = io.y1_pre; a.ldy = H; a.rows = N;
        MGB_TRY(launch_linear_tc(sh.precision, a, s));
        LinTcArgs b{};
        b.src[0] = io.y1_pre; b.ld[0] = H; b.nk = 1; b.self_act = ACT_SWISH;
        b.wimg = p.img_w4; b.nm = 1; b.tile_of[0][0] = 0;
        b.bias = io.b4; b.act = ACT_SWISH; b.residual = io.x; b.ldr = H;
        b.y = out; b.ldy = H; b.y_pre = io.y2_pre; b.ldyp = H; b.rows = N;
        MGB_TRY(launch_linear_tc(sh.precision, b, s));
    } else {
    {
        GemmArgs g{};
        g.a.p[0] = io.x; g.a.ld[0] = H; g.a.k[0] = H;
        g.a.p[1] = io.agg; g.a.ld[1] = H; g.a.k[1] = H;
        g.a.p[2] = io.var; g.a.ld[2] = sh.nv; g.a.k[2] = sh.nv;
        g.a.nseg = 3;
        g.b = p.w3t; g.ldb = H; g.bias = io.b3;
        g.c = y1; g.ldc = H; g.c_pre = io.y1_pre; g.ldcp = H; g.act = ACT_SWISH;
        g.M = N; g.N = H; g.K = sh.K3();
        MGB_TRY(launch_gemm(g, s));
    }
    {
        GemmArgs g{};
        g.a.p[0] = y1; g.a.ld[0] = H; g.a.k[0] = H; g.a.nseg = 1;
        g.b = p.w4t; g.ldb = H; g.bias = io.b4;
        g.residual = io.x; g.ldr = H;
        g.c = out; g.ldc = H; g.c_pre = io.y2_pre; g.ldcp = H; g.act = ACT_SWISH;
        g.M = N; g.N = H; g.K = H;
        MGB_TRY(launch_gemm(g, s));
    }
    }
    // 4. InstanceNorm per graph
    MGB_TRY(instance_norm_fwd(out, io.gptr, sh.n_graphs, sh.max_nodes_per_graph, io.y, io.rstd, ws.base + ws.off,
                              ws.cap - ws.off, s));
    return MGB_OK;
}

__global__ void __launch_bounds__(256)
gnn_combine_grads_kernel(const float* __restrict__ d0, const float* __restrict__ dc, int ld_dc, const float* __restrict__ dxc,
                         int ld_dxc, int64_t n, int tw, int dp, int nv, float* __restrict__ dx, float* __restrict__ du,
                         float* __restrict__ dpos, float* __restrict__ dvar) {
    const int Kc = H + tw + dp + nv;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * Kc) return;
    const int64_t r = i / Kc;
    const int k = (int)(i - r * Kc);
    const float v = dxc[r * ld_dxc + k];
    if (k < H) dx[r * H + k] = d0[r * H + k] + dc[r * ld_dc + k] + v;
    else if (k < H + tw) { if (du) du[r * tw + (k - H)] = v; }
    else if (k < H + tw + dp) { if (dpos) dpos[r * dp + (k - H - tw)] = v; }
    else if (dvar) dvar[r * nv + (k - H - tw - dp)] = v + dc[r * ld_dc + 2 * H + (k - H - tw - dp)];
}

static int ld4(int k) { return (k + 3) / 4 * 4; }

size_t gnn_layer_bwd_workspace(int64_t n_nodes, int64_t n_edges, int tw, int dp, int nv, int n_graphs, int max_nodes) {
    GnnLayerShape sh{n_nodes, n_edges, tw, dp, nv, n_graphs, max_nodes, 0};
    const int64_t tiles = ceil_div<int64_t>(n_edges > 0 ? n_edges : 1, TE);
    const int grid = edge_grid(tiles, 1);
    size_t b = 0;
    b += 2 * align_up((size_t)n_nodes * H * 4);                       // d0, d1
    b += align_up((size_t)n_nodes * ld4(sh.K3()) * 4);                // dc
    b += align_up((size_t)n_nodes * 2 * H * 4);                       // dpq
    b += align_up((size_t)n_nodes * ld4(sh.Kc()) * 4);                // dxc
    b += align_up((size_t)((n_edges > 0 ? n_edges : 1) + 127) / 128 * 128 * H * 4);       // dz1, padded to whole edge tiles
    b += 2 * align_up((size_t)tiles * H * 4);                         // part head/tail
    b += align_up((size_t)grid * H * H * 4) + align_up((size_t)grid * H * 4);
    b += align_up((size_t)2 * H * sh.Kc() * 4) + align_up(2 * H * 4);  // dwcat, dbcat
    size_t w = wgrad_workspace_bytes((int)n_nodes, 2 * H, sh.Kc());
    size_t w2 = wgrad_workspace_bytes((int)n_nodes, H, sh.K3());
    size_t w3 = wgrad_workspace_bytes((int)n_nodes, H, H);
    if (w2 > w) w = w2;
    if (w3 > w) w = w3;
    b += w + inorm_workspace_bytes(n_graphs, max_nodes) + 8192;
    b += align_up(edge_bwd_tc_workspace(n_edges)) + align_up(wgrad_tc_workspace(n_nodes));
    return b;
}

int gnn_layer_bwd(const GnnLayerShape& sh, const GnnBwdIO& io, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(sh.n_nodes < ((int64_t)1 << 31) && sh.n_edges < ((int64_t)1 << 31), "gnn_layer: sizes out of range");
    const int N = (int)sh.n_nodes;
    if (N == 0) return MGB_OK;
    GnnPacked p;
    packed_layout(sh, const_cast<float*>(io.packed), &p);
    Workspace ws(ws_ptr, ws_bytes);
    const int64_t tiles = ceil_div<int64_t>(sh.n_edges > 0 ? sh.n_edges : 1, TE);
    const int grid = edge_grid(tiles, 1);
    const int ldc = ld4(sh.K3()), ldx = ld4(sh.Kc());
    float* d0 = ws.take<float>((size_t)N * H);
    float* d1 = ws.take<float>((size_t)N * H);
    float* dc = ws.take<float>((size_t)N * ldc);
    float* dpq = ws.take<float>((size_t)N * 2 * H);
    float* dxc = ws.take<float>((size_t)N * ldx);
    float* dz1 = ws.take<float>((size_t)((sh.n_edges > 0 ? sh.n_edges : 1) + 127) / 128 * 128 * H);   // padded to whole edge tiles
    float* part_head = ws.take<float>((size_t)tiles * H);
    float* part_tail = ws.take<float>((size_t)tiles * H);
    float* dw2_part = ws.take<float>((size_t)grid * H * H);
    float* db2_part = ws.take<float>((size_t)grid * H);
    float* dwcat = ws.take<float>((size_t)2 * H * sh.Kc());
    float* dbcat = ws.take<float>(2 * H);
    const size_t tc_bytes = edge_bwd_tc_workspace(sh.n_edges);
    char* tc_ws = ws.take<char>(tc_bytes);
    const size_t wg_bytes = wgrad_tc_workspace(N);
    char* wg_ws = ws.take<char>(wg_bytes);
    MGB_WS_CHECK(ws);
    MGB_REQUIRE(sh.precision >= 0 && sh.precision <= 2, "gnn_layer: unknown precision %d", sh.precision);
    const int kt_pq = sh.tw + sh.dp + sh.nv;
    const bool tc_nodes = sh.precision != 0 && kt_pq <= 16 && sh.nv <= 16;
    void* sub_ws = ws.base + ws.off;
    size_t sub_bytes = ws.cap - ws.off;
    const int acc = io.accumulate_params;
    // nobody asked for du / dpos / dvar (the usual case: they are data): dx comes straight out of the last data-gradient Linear
    const bool fuse_dx = tc_nodes && !io.du && !io.dpos && !io.dvar;

    // B1. InstanceNorm backward -> d0 = d(out)
    MGB_TRY(instance_norm_bwd(io.dy, io.y, io.rstd, io.gptr, sh.n_graphs, sh.max_nodes_per_graph, d0, sub_ws, sub_bytes, s));
    // B2. update_net_2:  dW4 = (d0*Swish'(y2_pre))^T Swish(y1_pre);  d1 = (d0*Swish'(y2_pre)) W4
    if (tc_nodes) {
        WgradTcArgs w{};
        w.dy = d0; w.lddy = H; w.ny = 1; w.y_pre = io.y2_pre; w.ldyp = H; w.y_act = ACT_SWISH;
        w.nx = 1; w.x[0] = io.y1_pre; w.ldx[0] = H; w.x_act[0] = ACT_SWISH; w.rows = N;
        w.db[0] = io.db4; w.db_accumulate = acc;
        WgradTcOut o[1] = {};
        o[0].dw = io.dW4; o[0].lddw = H; o[0].n_valid = H; o[0].k_valid = H; o[0].accumulate = acc;
        MGB_TRY(launch_wgrad_tc(sh.precision, w, o, wg_ws, wg_bytes, s));
        LinTcArgs a{};
        a.src[0] = d0; a.ld[0] = H; a.nk = 1; a.pre = io.y2_pre; a.ldpre = H; a.pre_act = ACT_SWISH;
        a.wimg = p.img_w4; a.nm = 1; a.a_trans = 1; a.tile_of[0][0] = 0;
        a.act = ACT_NONE; a.y = d1; a.ldy = H; a.rows = N;
        if (ts_linears()) { a.wimg = p.wtm_d1; a.a_trans = 0; MGB_TRY(launch_linear_ts(sh.precision, a, s)); }
        else MGB_TRY(launch_linear_tc(sh.precision, a, s));
    } else {
        WgradArgs w{};
        w.dy = d0; w.lddy = H; w.y_pre = io.y2_pre; w.y_act = ACT_SWISH;
        w.a.p[0] = io.y1_pre; w.a.ld[0] = H; w.a.k[0] = H; w.a.nseg = 1; w.a.self_act = ACT_SWISH;
        w.rows = N; w.N = H; w.K = H; w.dw = io.dW4; w.lddw = H; w.db = io.db4; w.accumulate = acc;
        MGB_TRY(launch_wgrad(w, sub_ws, sub_bytes, s));
        GemmArgs g{};
        g.a.p[0] = d0; g.a.ld[0] = H; g.a.k[0] = H; g.a.nseg = 1; g.a.pre = io.y2_pre; g.a.pre_act = ACT_SWISH;
        g.b = io.W4; g.ldb = H; g.c = d1; g.ldc = H; g.M = N; g.N = H; g.K = H;
        MGB_TRY(launch_gemm(g, s));
    }
    // B3. update_net_1:  dW3 = (d1*Swish'(y1_pre))^T [x,agg,var];  dc = (d1*Swish'(y1_pre)) W3
    if (tc_nodes) {
        WgradTcArgs w{};
        w.dy = d1; w.lddy = H; w.ny = 1; w.y_pre = io.y1_pre; w.ldyp = H; w.y_act = ACT_SWISH; w.rows = N;
        w.nx = 2; w.x[0] = io.x; w.ldx[0] = H; w.x[1] = io.agg; w.ldx[1] = H;
        w.tail = 1; w.tsrc[0] = io.var; w.tld[0] = sh.nv; w.tk[0] = sh.nv; w.kt = sh.nv;
        WgradTcOut o[3] = {};
        o[0].dw = io.dW3; o[1].dw = io.dW3 + H; o[2].dw = io.dW3 + 2 * H;
        for (int i = 0; i < 3; ++i) { o[i].lddw = sh.K3(); o[i].n_valid = H; o[i].k_valid = i < 2 ? H : sh.nv; o[i].accumulate = acc; }
        w.db[0] = io.db3; w.db_accumulate = acc;
        MGB_TRY(launch_wgrad_tc(sh.precision, w, o, wg_ws, wg_bytes, s));
        LinTcArgs a{};
        a.src[0] = d1; a.ld[0] = H; a.nk = 1; a.pre = io.y1_pre; a.ldpre = H; a.pre_act = ACT_SWISH;
        a.wimg = p.img_w3; a.nm = 2; a.a_trans = 1; a.tile_of[0][0] = 0; a.tile_of[1][0] = 1;
        a.act = ACT_NONE; a.y = dc; a.ldy = ldc; a.rows = N;
        if (fuse_dx) { a.residual = d0; a.ldr = H; a.res_blocks = 1; }      // dc[:, :H] += d0 (the residual branch of the update)
        if (ts_linears()) { a.wimg = p.wtm_dc; a.a_trans = 0; MGB_TRY(launch_linear_ts(sh.precision, a, s)); }
        else MGB_TRY(launch_linear_tc(sh.precision, a, s));
        if (io.dvar)      // the tail columns of dc feed dvar only
            MGB_TRY(launch_tail_dgrad(d1, H, H, io.y1_pre, H, ACT_SWISH, io.W3 + 2 * H, sh.K3(), 1, sh.nv, N, dc + 2 * H, ldc, s));
    } else {
        WgradArgs w{};
        w.dy = d1; w.lddy = H; w.y_pre = io.y1_pre; w.y_act = ACT_SWISH;
        w.a.p[0] = io.x; w.a.ld[0] = H; w.a.k[0] = H;
        w.a.p[1] = io.agg; w.a.ld[1] = H; w.a.k[1] = H;
        w.a.p[2] = io.var; w.a.ld[2] = sh.nv; w.a.k[2] = sh.nv; w.a.nseg = 3;
        w.rows = N; w.N = H; w.K = sh.K3(); w.dw = io.dW3; w.lddw = sh.K3(); w.db = io.db3; w.accumulate = acc;
        MGB_TRY(launch_wgrad(w, sub_ws, sub_bytes, s));
        GemmArgs g{};
        g.a.p[0] = d1; g.a.ld[0] = H; g.a.k[0] = H; g.a.nseg = 1; g.a.pre = io.y1_pre; g.a.pre_act = ACT_SWISH;
        g.b = io.W3; g.ldb = sh.K3(); g.c = dc; g.ldc = ldc; g.M = N; g.N = sh.K3(); g.K = H;
        MGB_TRY(launch_gemm(g, s));
    }
    // B4. fused edge backward (dagg = dc[:, H:2H])
    if (sh.n_edges > 0 && sh.precision != 0) {      // dP: rows of nodes with in-edges are stored in full; dQ: gather_sum_rows writes every row
        zero_isolated_rows_kernel<<<(unsigned)ceil_div<int64_t>(N, 256), 256, 0, s>>>(io.rowptr, N, dpq, 2 * H);
        MGB_LAUNCH_CHECK();
    } else {
        MGB_CUDA(cudaMemsetAsync(dpq, 0, (size_t)N * 2 * H * sizeof(float), s));
    }
    if (sh.n_edges > 0 && sh.precision != 0) {
        MGB_TRY(launch_edge_bwd_tc(sh.precision, io.pq, io.rowptr, io.dstv, io.srcv, sh.n_edges, p.w2img, io.b2, dc + H, ldc,
                                   dz1, dpq, io.dW2, io.db2, acc, tc_ws, tc_bytes, s));
        gather_sum_rows_kernel<<<(unsigned)ceil_div<int64_t>((int64_t)N * 32, 256), 256, 0, s>>>(dz1, io.rowptr_t, io.pos_t, N, dpq + H, 2 * H);
        MGB_LAUNCH_CHECK();
    } else if (sh.n_edges > 0) {
        EdgeBwdArgs a{io.pq, io.rowptr, io.dstv, io.srcv, sh.n_edges, p.w2t, io.W2, io.b2, dc + H, ldc,
                      dz1, dpq, part_head, part_tail, dw2_part, db2_part};
        MGB_CUDA(cudaFuncSetAttribute(gnn_edge_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EDGE_BWD_SMEM));
        {
            ProfScope prof(PROF_EDGE_BWD, s);
            gnn_edge_bwd_kernel<<<grid, 256, EDGE_BWD_SMEM, s>>>(a);
        }
        MGB_LAUNCH_CHECK();
        if (tiles > 1) {
            segment_fixup_kernel<false><<<(unsigned)(tiles - 1), H, 0, s>>>(io.rowptr, io.dstv, sh.n_edges, part_head, part_tail, dpq, 2 * H);
            MGB_LAUNCH_CHECK();
        }
        sum_partials_kernel<<<ceil_div(H * H, 256), 256, 0, s>>>(dw2_part, grid, (int64_t)H * H, io.dW2, acc);
        MGB_LAUNCH_CHECK();
        sum_partials_kernel<<<1, 256, 0, s>>>(db2_part, grid, H, io.db2, acc);
        MGB_LAUNCH_CHECK();
        // B5. dQ[j] = sum of dz1 over the edges leaving j (transposed CSR)
        gather_sum_rows_kernel<<<(unsigned)ceil_div<int64_t>((int64_t)N * 32, 256), 256, 0, s>>>(dz1, io.rowptr_t, io.pos_t, N, dpq + H, 2 * H);
        MGB_LAUNCH_CHECK();
    } else if (!acc) {
        MGB_CUDA(cudaMemsetAsync(io.dW2, 0, (size_t)H * H * sizeof(float), s));
        MGB_CUDA(cudaMemsetAsync(io.db2, 0, (size_t)H * sizeof(float), s));
    }
    // B6. first Linear (factorised): dWcat = dPQ^T [x,u,pos,var];  dxc = dPQ Wcat
    if (tc_nodes) {
        {   // rows of dWcat: P outputs (Y' tile 0), Q outputs (Y' tile 1); columns: x | u,pos,var; dbcat from the ones column
            WgradTcArgs w{};
            w.dy = dpq; w.lddy = 2 * H; w.ny = 2; w.rows = N;
            w.nx = 1; w.x[0] = io.x; w.ldx[0] = H;
            w.tail = 1;
            w.tsrc[0] = io.u; w.tld[0] = sh.tw; w.tk[0] = sh.tw;
            w.tsrc[1] = io.pos; w.tld[1] = sh.dp; w.tk[1] = sh.dp;
            w.tsrc[2] = io.var; w.tld[2] = sh.nv; w.tk[2] = sh.nv; w.kt = kt_pq;
            WgradTcOut o[4] = {};
            for (int half = 0; half < 2; ++half) {
                WgradTcOut& ox = o[half * 2];
                WgradTcOut& ot = o[half * 2 + 1];
                ox.dw = dwcat + (size_t)half * H * sh.Kc(); ox.lddw = sh.Kc(); ox.n_valid = H; ox.k_valid = H;
                ot.dw = dwcat + (size_t)half * H * sh.Kc() + H; ot.lddw = sh.Kc(); ot.n_valid = H; ot.k_valid = kt_pq;
                w.db[half] = dbcat + half * H;
            }
            MGB_TRY(launch_wgrad_tc(sh.precision, w, o, wg_ws, wg_bytes, s));
        }
        unpack_dw1_kernel<<<ceil_div(H * sh.K1(), 256), 256, 0, s>>>(dwcat, sh.tw, sh.dp, sh.nv, io.dW1, acc);
        MGB_LAUNCH_CHECK();
        sum_partials_kernel<<<1, 256, 0, s>>>(dbcat, 1, H, io.db1, acc);
        MGB_LAUNCH_CHECK();
        LinTcArgs a{};
        a.src[0] = dpq; a.ld[0] = 2 * H; a.src[1] = dpq + H; a.ld[1] = 2 * H; a.nk = 2;
        a.wimg = p.img_pq; a.nm = 1; a.a_trans = 1; a.tile_of[0][0] = 0; a.tile_of[0][1] = 1;
        a.act = ACT_NONE; a.y = dxc; a.ldy = ldx; a.rows = N;
        if (fuse_dx) { a.y = io.dx; a.ldy = H; a.residual = dc; a.ldr = ldc; }   // dx = dPQ Wcat[:, :H] + (d0 + dc[:, :H]): no assembly pass
        if (ts_linears()) { a.wimg = p.wtm_dx; a.a_trans = 0; MGB_TRY(launch_linear_ts(sh.precision, a, s)); }
        else MGB_TRY(launch_linear_tc(sh.precision, a, s));
        if (io.du || io.dpos || io.dvar)      // the tail columns of dxc feed du / dpos / dvar only
            MGB_TRY(launch_tail_dgrad(dpq, 2 * H, 2 * H, nullptr, 0, ACT_NONE, p.wcat + H, sh.Kc(), 1, kt_pq, N, dxc + H, ldx, s));
    } else {
        WgradArgs w{};
        w.dy = dpq; w.lddy = 2 * H;
        w.a.p[0] = io.x; w.a.ld[0] = H; w.a.k[0] = H;
        w.a.p[1] = io.u; w.a.ld[1] = sh.tw; w.a.k[1] = sh.tw;
        w.a.p[2] = io.pos; w.a.ld[2] = sh.dp; w.a.k[2] = sh.dp;
        w.a.p[3] = io.var; w.a.ld[3] = sh.nv; w.a.k[3] = sh.nv; w.a.nseg = 4;
        w.rows = N; w.N = 2 * H; w.K = sh.Kc(); w.dw = dwcat; w.lddw = sh.Kc(); w.db = dbcat; w.accumulate = 0;
        MGB_TRY(launch_wgrad(w, sub_ws, sub_bytes, s));
        unpack_dw1_kernel<<<ceil_div(H * sh.K1(), 256), 256, 0, s>>>(dwcat, sh.tw, sh.dp, sh.nv, io.dW1, acc);
        MGB_LAUNCH_CHECK();
        sum_partials_kernel<<<1, 256, 0, s>>>(dbcat, 1, H, io.db1, acc);
        MGB_LAUNCH_CHECK();
        GemmArgs g{};
        g.a.p[0] = dpq; g.a.ld[0] = 2 * H; g.a.k[0] = 2 * H; g.a.nseg = 1;
        g.b = p.wcat; g.ldb = sh.Kc(); g.c = dxc; g.ldc = ldx; g.M = N; g.N = sh.Kc(); g.K = 2 * H;
        MGB_TRY(launch_gemm(g, s));
    }
    // B7. assemble the input gradients
    if (!fuse_dx)
    gnn_combine_grads_kernel<<<(unsigned)ceil_div<int64_t>((int64_t)N * sh.Kc(), 256), 256, 0, s>>>(
        d0, dc, ldc, dxc, ldx, N, sh.tw, sh.dp, sh.nv, io.dx, io.du, io.dpos, io.dvar);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
