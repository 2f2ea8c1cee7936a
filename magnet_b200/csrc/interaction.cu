// magnet_b200 — MAgNet[GNN] specific stages: InteractionNetwork edge/aggregation glue and the INR decoder.
//
// Reference:
//   InteractionNetwork.message / aggregate  models/magnet_gnn.py:70-90 (x_i = x[edge_index[1]], x_j = x[edge_index[0]],
//       mean aggregation at edge_index[1], aggr='mean' :54)
//   MAgNetGNN.continuous_decoder           models/magnet_gnn.py:224-283
//
// The first Linear of edge_fn is factorised (W [x_i, x_j, e] = P[dst] + Q[src] + R[e]); the kernels here combine
// the three terms per edge, reduce messages per destination through the sorted plan (no atomics), and evaluate the
// decoder per query point with the latent part of proj_head factorised per low-resolution node
// (A = enc Wp[:, :C]^T + b once per node; the k-2 neighbours the reference computes and then ignores, SURVEY F9, are skipped).
#include "internal.cuh"

namespace mgb {

constexpr int IH = 128;

// out[e] = act(P[i_e] + Q[j_e] + R[e]),  i = edge_index[1] (aggregation endpoint), j = edge_index[0].  One warp per edge.
__global__ void __launch_bounds__(256)
edge_combine_fwd_kernel(const float* __restrict__ P, const float* __restrict__ Q, const float* __restrict__ R,
                        const int64_t* __restrict__ ei0, const int64_t* __restrict__ ei1, int64_t n_edges, int act,
                        float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= n_edges) return;
    const int64_t i = ei1[e], j = ei0[e];
    const float4 p = *reinterpret_cast<const float4*>(P + i * IH + lane * 4);
    const float4 q = *reinterpret_cast<const float4*>(Q + j * IH + lane * 4);
    const float4 r = *reinterpret_cast<const float4*>(R + e * IH + lane * 4);
    float4 o = make_float4(p.x + q.x + r.x, p.y + q.y + r.y, p.z + q.z + r.z, p.w + q.w + r.w);
    o.x = act_apply(act, o.x); o.y = act_apply(act, o.y); o.z = act_apply(act, o.z); o.w = act_apply(act, o.w);
    *reinterpret_cast<float4*>(out + e * IH + lane * 4) = o;
}

// dz = dout * (out > 0)   (ReLU backward from the saved OUTPUT)
__global__ void __launch_bounds__(256)
relu_mask_kernel(const float* __restrict__ dout, const float* __restrict__ out, int64_t n4, float* __restrict__ dz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 g = reinterpret_cast<const float4*>(dout)[i], o = reinterpret_cast<const float4*>(out)[i];
    reinterpret_cast<float4*>(dz)[i] = make_float4(o.x > 0.f ? g.x : 0.f, o.y > 0.f ? g.y : 0.f, o.z > 0.f ? g.z : 0.f, o.w > 0.f ? g.w : 0.f);
}

// out[n][c] = scale * sum_{q in [rowptr[n], rowptr[n+1])} rows[idx ? idx[q] : q][c],  scale = mean ? 1/max(count,1) : 1.
// One warp per node; fixed summation order (deterministic, no atomics).
__global__ void __launch_bounds__(256)
segment_sum_rows_kernel(const float* __restrict__ rows, int cols, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ idx,
                        int64_t n_nodes, int mean, float* __restrict__ out, int ld_out) {
    const int lane = threadIdx.x & 31;
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= n_nodes) return;
    const int s = rowptr[n], e = rowptr[n + 1];
    const float scale = mean ? 1.0f / (float)(e - s > 1 ? e - s : 1) : 1.0f;
    if (cols == IH) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = s; q < e; ++q) {
            const int64_t r = idx ? idx[q] : q;
            const float4 v = *reinterpret_cast<const float4*>(rows + r * IH + lane * 4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
        *reinterpret_cast<float4*>(out + n * ld_out + lane * 4) = acc;
    } else {
        for (int c = lane; c < cols; c += 32) {
            float acc = 0.f;
            for (int q = s; q < e; ++q) acc += rows[(int64_t)(idx ? idx[q] : q) * cols + c];
            out[n * ld_out + c] = acc * scale;
        }
    }
}

// out[e] = rows[index[e]] * (rowptr ? 1/max(count(index[e]),1) : 1)       (backward of the mean aggregation)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ rows, const int64_t* __restrict__ index, const int32_t* __restrict__ rowptr,
                   int64_t n_edges, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= n_edges) return;
    const int64_t i = index[e];
    float scale = 1.0f;
    if (rowptr) {
        const int c = rowptr[i + 1] - rowptr[i];
        scale = 1.0f / (float)(c > 1 ? c : 1);
    }
    float4 v = *reinterpret_cast<const float4*>(rows + i * IH + lane * 4);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4*>(out + e * IH + lane * 4) = v;
}

int edge_combine_fwd(const float* P, const float* Q, const float* R, const int64_t* edge_index, int64_t n_edges, int act,
                     float* out, cudaStream_t s) {
    if (n_edges <= 0) return MGB_OK;
    MGB_REQUIRE(act == ACT_NONE || act == ACT_RELU, "edge_combine: act must be none or relu");
    edge_combine_fwd_kernel<<<(unsigned)ceil_div<int64_t>(n_edges * 32, 256), 256, 0, s>>>(P, Q, R, edge_index, edge_index + n_edges,
                                                                                             n_edges, act, out);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

int relu_mask(const float* dout, const float* out, int64_t n, float* dz, cudaStream_t s) {
    if (n <= 0) return MGB_OK;
    MGB_REQUIRE(n % 4 == 0, "relu_mask: element count must be a multiple of 4");
    relu_mask_kernel<<<(unsigned)ceil_div<int64_t>(n / 4, 256), 256, 0, s>>>(dout, out, n / 4, dz);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

int segment_sum_rows(const float* rows, int cols, const int32_t* rowptr, const int32_t* idx, int64_t n_nodes, int mean,
                     float* out, int ld_out, cudaStream_t s) {
    if (n_nodes <= 0) return MGB_OK;
    MGB_REQUIRE(cols >= 1 && ld_out >= cols, "segment_sum_rows: bad cols/ld_out");
    segment_sum_rows_kernel<<<(unsigned)ceil_div<int64_t>(n_nodes * 32, 256), 256, 0, s>>>(rows, cols, rowptr, idx, n_nodes, mean, out, ld_out);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

int gather_rows(const float* rows, const int64_t* index, const int32_t* rowptr, int64_t n_edges, float* out, cudaStream_t s) {
    if (n_edges <= 0) return MGB_OK;
    gather_rows_kernel<<<(unsigned)ceil_div<int64_t>(n_edges * 32, 256), 256, 0, s>>>(rows, index, rowptr, n_edges, out);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------------------
// INR decoder (continuous_decoder).  One warp per query point, 4 channels per lane.
//   lat_j(i)[c] = A[sel_j][c] + wu[c] x_lr[b,i,sel_j] + sum_d wrel[c][d] rel_j[d] + wt[c] t[b,i]
//   rel_j = lr_coords[sel_j] - hr_coords[q],  s_j = (||rel_j||_2)^2
//   area: z = (lat_0 s_1 + lat_1 s_0) / (s_1 + s_0);  knn: w = 1/s;  sph: w = (1 - L s)^3;  z = (lat_0 w_0 + lat_1 w_1)/(w_1 + w_0)
// wsmall points at column C of proj_head.weight (row stride ldw = C + d + 2): columns [u, rel_0..rel_{d-1}, t].
// ------------------------------------------------------------------------------------------------------

// m0 / m1: the weights that multiply lat_0 / lat_1 in the reference's blend, den = w_1 + w_0; a_j = m_j / den
__device__ __forceinline__ void inr_weights(const InrArgs& a, int64_t q, int64_t& sel0, int64_t& sel1, float (&rel0)[2], float (&rel1)[2],
                                            float& a0, float& a1, float& m0, float& m1, float& den) {
    sel0 = a.idx[q * a.k];
    sel1 = a.idx[q * a.k + 1];
    float s0 = 0.f, s1 = 0.f;
    rel0[1] = rel1[1] = 0.f;
    for (int dd = 0; dd < a.d; ++dd) {
        const float h = a.hr_coords[q * a.d + dd];
        rel0[dd] = a.lr_coords[sel0 * a.d + dd] - h;
        rel1[dd] = a.lr_coords[sel1 * a.d + dd] - h;
        s0 += rel0[dd] * rel0[dd];
        s1 += rel1[dd] * rel1[dd];
    }
    float n0 = sqrtf(s0), n1 = sqrtf(s1);      // the reference squares torch.norm(...)
    float w0 = n0 * n0, w1 = n1 * n1;
    if (a.mode == 1) { w0 = 1.0f / w0; w1 = 1.0f / w1; }
    if (a.mode == 2) { const float b0 = 1.0f - (float)a.L * w0, b1 = 1.0f - (float)a.L * w1; w0 = b0 * b0 * b0; w1 = b1 * b1 * b1; }
    den = w1 + w0;
    if (a.mode == 0) { m0 = w1; m1 = w0; }     // area: neighbour 0 is weighted by the OTHER one's area
    else { m0 = w0; m1 = w1; }
    a0 = m0 / den;
    a1 = m1 / den;
}

__global__ void __launch_bounds__(256) inr_decode_fwd_kernel(const InrArgs a, float* __restrict__ z) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= a.n_query) return;
    int64_t sel0, sel1;
    float rel0[2], rel1[2], a0, a1, m0, m1, den;
    inr_weights(a, q, sel0, sel1, rel0, rel1, a0, a1, m0, m1, den);
    const int b = (int)(q / a.nq_per_sample);
    const int c = lane * 4;
    float wu[4], wt[4], c0[4], c1[4];
    const float4 A0 = *reinterpret_cast<const float4*>(a.A + sel0 * IH + c);
    const float4 A1 = *reinterpret_cast<const float4*>(a.A + sel1 * IH + c);
    const float A0v[4] = {A0.x, A0.y, A0.z, A0.w}, A1v[4] = {A1.x, A1.y, A1.z, A1.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float* w = a.wsmall + (int64_t)(c + u) * a.ldw;
        wu[u] = w[0];
        wt[u] = w[1 + a.d];
        float r0 = A0v[u], r1 = A1v[u];
        for (int dd = 0; dd < a.d; ++dd) { r0 = fmaf(w[1 + dd], rel0[dd], r0); r1 = fmaf(w[1 + dd], rel1[dd], r1); }
        c0[u] = r0; c1[u] = r1;       // time-independent part of lat_0 / lat_1
    }
    const int64_t l0 = sel0 - (int64_t)b * a.L, l1 = sel1 - (int64_t)b * a.L;
    for (int i = 0; i < a.T; ++i) {
        const float x0 = a.xlr[((int64_t)b * a.T + i) * a.L + l0], x1 = a.xlr[((int64_t)b * a.T + i) * a.L + l1];
        const float ti = a.t[(int64_t)b * a.ldt + i];
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float lat0 = fmaf(wt[u], ti, fmaf(wu[u], x0, c0[u]));
            const float lat1 = fmaf(wt[u], ti, fmaf(wu[u], x1, c1[u]));
            // same operation order as the reference: (lat_0 m_0 + lat_1 m_1) / (w_1 + w_0), no contraction
            o[u] = __fdiv_rn(__fadd_rn(__fmul_rn(lat0, m0), __fmul_rn(lat1, m1)), den);
        }
        *reinterpret_cast<float4*>(z + (q * a.T + i) * IH + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// Backward w.r.t. A (through G, two rows per query), x_lr (through Sx) and the small proj_head columns.
//   G[2q+j][c] = a_j sum_i dz[q][i][c];  Sx[2q+j][i] = a_j sum_c dz[q][i][c] wu[c]
//   dwsmall partial per warp: [128][d+2] accumulated over the warp's queries (grid-stride), reduced by a second stage.
__global__ void __launch_bounds__(256)
inr_decode_bwd_kernel(const InrArgs a, const float* __restrict__ dz, float* __restrict__ G, float* __restrict__ Sx,
                      float* __restrict__ dw_partial, int n_warps_total) {
    const int lane = threadIdx.x & 31;
    const int warp_id = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int c = lane * 4;
    const int ncol = a.d + 2;
    float wu[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wu[u] = a.wsmall[(int64_t)(c + u) * a.ldw];
    float dw[4][4];       // [channel u][column: u, rel0, rel1, t]  (d <= 2)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) dw[u][v] = 0.f;
    for (int64_t q = warp_id; q < a.n_query; q += n_warps_total) {
        int64_t sel0, sel1;
        float rel0[2], rel1[2], a0, a1, m0, m1, den;
        inr_weights(a, q, sel0, sel1, rel0, rel1, a0, a1, m0, m1, den);
        const int b = (int)(q / a.nq_per_sample);
        const int64_t l0 = sel0 - (int64_t)b * a.L, l1 = sel1 - (int64_t)b * a.L;
        const float rb0 = a0 * rel0[0] + a1 * rel1[0], rb1 = a0 * rel0[1] + a1 * rel1[1];
        float gs[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < a.T; ++i) {
            const float4 g4 = *reinterpret_cast<const float4*>(dz + (q * a.T + i) * IH + c);
            const float g[4] = {g4.x, g4.y, g4.z, g4.w};
            const float x0 = a.xlr[((int64_t)b * a.T + i) * a.L + l0], x1 = a.xlr[((int64_t)b * a.T + i) * a.L + l1];
            const float ti = a.t[(int64_t)b * a.ldt + i];
            const float xb = a0 * x0 + a1 * x1;
            float dot = 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                gs[u] += g[u];
                dot = fmaf(g[u], wu[u], dot);
                dw[u][0] = fmaf(g[u], xb, dw[u][0]);
                dw[u][3] = fmaf(g[u], ti, dw[u][3]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            if (lane == 0) {
                Sx[(2 * q) * a.T + i] = a0 * dot;
                Sx[(2 * q + 1) * a.T + i] = a1 * dot;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dw[u][1] = fmaf(gs[u], rb0, dw[u][1]);
            dw[u][2] = fmaf(gs[u], rb1, dw[u][2]);
        }
        *reinterpret_cast<float4*>(G + (2 * q) * IH + c) = make_float4(a0 * gs[0], a0 * gs[1], a0 * gs[2], a0 * gs[3]);
        *reinterpret_cast<float4*>(G + (2 * q + 1) * IH + c) = make_float4(a1 * gs[0], a1 * gs[1], a1 * gs[2], a1 * gs[3]);
    }
    float* out = dw_partial + (int64_t)warp_id * IH * 4;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        out[(c + u) * 4 + 0] = dw[u][0];
        out[(c + u) * 4 + 1] = dw[u][1];
        out[(c + u) * 4 + 2] = ncol > 3 ? dw[u][2] : 0.f;
        out[(c + u) * 4 + 3] = dw[u][3];
    }
}

// dwsmall[c][col] (+)= sum over warps; partial layout [warps][128][4] = (u, rel0, rel1, t)
__global__ void __launch_bounds__(128)
inr_dw_reduce_kernel(const float* __restrict__ partial, int n_warps, int d, float* __restrict__ dw, int ldw, int accumulate) {
    const int c = threadIdx.x;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < n_warps; ++w) {
        const float4 v = *reinterpret_cast<const float4*>(partial + ((int64_t)w * IH + c) * 4);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
    float* o = dw + (int64_t)c * ldw;      // columns: u, rel_0, [rel_1,] t
    const float base[4] = {accumulate ? o[0] : 0.f, accumulate ? o[1] : 0.f, (accumulate && d > 1) ? o[2] : 0.f,
                           accumulate ? o[d + 1] : 0.f};
    o[0] = base[0] + s[0];
    o[1] = base[1] + s[1];
    if (d > 1) o[2] = base[2] + s[2];
    o[d + 1] = base[3] + s[3];
}

static int inr_check(const InrArgs& a) {
    MGB_REQUIRE(a.d == 1 || a.d == 2, "inr_decode: coordinate dimension must be 1 or 2");
    MGB_REQUIRE(a.k >= 2, "inr_decode: the decoder blends the two nearest nodes, k must be >= 2");
    MGB_REQUIRE(a.mode >= 0 && a.mode <= 2, "inr_decode: unknown interpolation mode");
    MGB_REQUIRE(a.T >= 1 && a.L >= 2 && a.nq_per_sample >= 1, "inr_decode: bad sizes");
    return MGB_OK;
}

int inr_decode_fwd(const InrArgs& a, float* z, cudaStream_t s) {
    MGB_TRY(inr_check(a));
    if (a.n_query <= 0) return MGB_OK;
    ProfScope prof(PROF_INR_DECODE, s);
    inr_decode_fwd_kernel<<<(unsigned)ceil_div<int64_t>(a.n_query * 32, 256), 256, 0, s>>>(a, z);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

static int inr_bwd_warps(int64_t n_query) {
    int64_t w = (int64_t)sm_count() * 16;
    if (w > n_query) w = n_query > 0 ? n_query : 1;
    return (int)((w + 7) / 8 * 8);
}

size_t inr_decode_bwd_workspace(int64_t n_query) { return align_up((size_t)inr_bwd_warps(n_query) * IH * 4 * sizeof(float)) + 256; }

int inr_decode_bwd(const InrArgs& a, const float* dz, float* G, float* Sx, float* dwsmall, int ldw, int accumulate, void* ws_ptr,
                   size_t ws_bytes, cudaStream_t s) {
    MGB_TRY(inr_check(a));
    const int warps = inr_bwd_warps(a.n_query);
    Workspace ws(ws_ptr, ws_bytes);
    float* partial = ws.take<float>((size_t)warps * IH * 4);
    MGB_WS_CHECK(ws);
    inr_decode_bwd_kernel<<<warps / 8, 256, 0, s>>>(a, dz, G, Sx, partial, warps);
    MGB_LAUNCH_CHECK();
    inr_dw_reduce_kernel<<<1, 128, 0, s>>>(partial, warps, a.d, dwsmall, ldw, accumulate);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------
// MAgNetGNN._build_graph features (models/magnet_gnn.py:298-308) in one launch instead of ten:
//   node_features[r] = [u[r, :C], x[r, :d], t_last[r % B]]                    (the time column is TILED: quirk F7)
//   edge_features[e] = [u[s_e] - u[r_e], x[s_e] - x[r_e]]                     s = edge_index[0], r = edge_index[1]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) magnet_features_kernel(const float* __restrict__ u, int C, const float* __restrict__ x, int d,
                                                              const float* __restrict__ t_last, int B, int64_t n_nodes,
                                                              const int64_t* __restrict__ edge_index, int64_t n_edges,
                                                              float* __restrict__ nf, float* __restrict__ ef) {
    const int wn = C + d + 1, we = C + d;
    const int64_t total_n = n_nodes * wn, total = total_n + n_edges * we;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < total_n) {
            const int64_t r = i / wn;
            const int c = (int)(i - r * wn);
            nf[i] = c < C ? u[r * C + c] : (c < C + d ? x[r * d + (c - C)] : t_last[r % B]);
        } else {
            const int64_t j = i - total_n, e = j / we;
            const int c = (int)(j - e * we);
            const int64_t sn = edge_index[e], rc = edge_index[n_edges + e];
            ef[j] = c < C ? u[sn * C + c] - u[rc * C + c] : x[sn * d + (c - C)] - x[rc * d + (c - C)];
        }
    }
}

int magnet_features(const float* u, int C, const float* x, int d, const float* t_last, int B, int64_t n_nodes,
                    const int64_t* edge_index, int64_t n_edges, float* nf, float* ef, cudaStream_t s) {
    MGB_REQUIRE(C >= 1 && d >= 1 && B >= 1 && n_nodes >= 0 && n_edges >= 0, "magnet_features: bad sizes");
    const int64_t total = n_nodes * (C + d + 1) + n_edges * (C + d);
    if (total == 0) return MGB_OK;
    const int64_t nb = ceil_div<int64_t>(total, 256);
    const int blocks = (int)(nb < (int64_t)sm_count() * 16 ? nb : (int64_t)sm_count() * 16);
    magnet_features_kernel<<<blocks, 256, 0, s>>>(u, C, x, d, t_last, B, n_nodes, edge_index, n_edges, nf, ef);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
