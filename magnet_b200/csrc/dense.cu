// magnet_b200 — row-wise dense kernels (fp32 FFMA): GEMM with fused prologue/epilogue, weight
// gradient with deterministic split-row reduction, column sums, transpose, LayerNorm fwd/bwd.
#include "dense.cuh"

namespace mgb {

// ------------------------------------------------------------------------------------------
// operand fetch helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fetch_a(const ASpec& a, int64_t row, int k) {
    // caller guarantees row < rows and k < sum(a.k)
    float v;
    if (k < a.k[0]) {
        v = a.p[0][row * a.ld[0] + k];
        if (a.pre) v *= act_grad(a.pre_act, a.pre[row * a.ld[0] + k]);
    } else if (k < a.k[0] + a.k[1]) {
        v = a.p[1][row * a.ld[1] + (k - a.k[0])];
    } else if (k < a.k[0] + a.k[1] + a.k[2]) {
        v = a.p[2][row * a.ld[2] + (k - a.k[0] - a.k[1])];
    } else {
        v = a.p[3][row * a.ld[3] + (k - a.k[0] - a.k[1] - a.k[2])];
    }
    if (a.self_act) v = act_apply(a.self_act, v);
    return v;
}

// ------------------------------------------------------------------------------------------
// GEMM: 128x128x16 tiles, 256 threads, 8x8 register tile (2x2 groups of 4x4), register-staged
// double buffering.
// ------------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 128, GK = 16;

__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[2][GK][GM + 4];
    __shared__ __align__(16) float Bs[2][GK][GN];
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;
    const int64_t row0 = (int64_t)blockIdx.x * GM;      // rows on grid.x: millions of edge rows exceed grid.y's 65,535
    const int col0 = blockIdx.y * GN;
    const int ar = tid >> 1, ak0 = (tid & 1) * 8;      // A loader: row ar, k offsets ak0..ak0+7
    const int bk = tid >> 4, bc0 = (tid & 15) * 8;     // B loader: k row bk, cols bc0..bc0+7
    float ra[8], rb[8];
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto load_regs = [&](int kt) {
        const int64_t r = row0 + ar;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int k = kt * GK + ak0 + i;
            ra[i] = (r < g.M && k < g.K) ? fetch_a(g.a, r, k) : 0.f;
        }
        const int kb = kt * GK + bk;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int c = col0 + bc0 + j;
            rb[j] = (kb < g.K && c < g.N) ? g.b[(int64_t)kb * g.ldb + c] : 0.f;
        }
    };
    auto store_regs = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][ak0 + i][ar] = ra[i];
        *reinterpret_cast<float4*>(&Bs[buf][bk][bc0]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
        *reinterpret_cast<float4*>(&Bs[buf][bk][bc0 + 4]) = make_float4(rb[4], rb[5], rb[6], rb[7]);
    };

    const int nkt = ceil_div(g.K, GK);
    load_regs(0);
    store_regs(0);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) load_regs(kt + 1);
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + tm * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tn * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nkt) {
            store_regs(buf ^ 1);
            __syncthreads();
        }
    }
    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = row0 + (i < 4 ? tm * 4 + i : 64 + tm * 4 + (i - 4));
        if (r >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col0 + (j < 4 ? tn * 4 + j : 64 + tn * 4 + (j - 4));
            if (c >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += g.bias[c];
            if (g.accumulate) v += g.c[r * g.ldc + c];
            if (g.c_pre) g.c_pre[r * g.ldcp + c] = v;
            v = act_apply(g.act, v);
            if (g.residual) v += g.residual[r * g.ldr + c];
            g.c[r * g.ldc + c] = v;
        }
    }
}

int launch_gemm(const GemmArgs& g, cudaStream_t s) {
    if (g.M == 0 || g.N == 0) return MGB_OK;
    int ksum = 0;
    for (int i = 0; i < g.a.nseg; ++i) ksum += g.a.k[i];
    MGB_REQUIRE(g.a.nseg >= 1 && g.a.nseg <= 4 && ksum == g.K, "gemm: A segments (%d) do not add up to K=%d", ksum, g.K);
    MGB_REQUIRE(!(g.a.pre && g.a.nseg != 1), "gemm: activation-gradient prologue needs a single A segment");
    MGB_REQUIRE(!(g.accumulate && g.act != ACT_NONE), "gemm: accumulate with an activation is not supported");
    GemmArgs a = g;
    for (int i = g.a.nseg; i < 4; ++i) { a.a.p[i] = nullptr; a.a.ld[i] = 0; a.a.k[i] = 0; }
    dim3 grid((unsigned)ceil_div<int64_t>(g.M, GM), ceil_div(g.N, GN));
    gemm_kernel<<<grid, 256, 0, s>>>(a);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------
// weight gradient: split over row chunks, deterministic second-stage reduction
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs w, float* __restrict__ partial, int chunk_rows) {
    __shared__ __align__(16) float Ys[2][16][128];
    __shared__ __align__(16) float As[2][16][128];
    const int tid = threadIdx.x;
    const int tn = tid >> 4, tk = tid & 15;
    const int k0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
    const int64_t rbeg = (int64_t)blockIdx.z * chunk_rows;
    const int64_t rend = min((int64_t)w.rows, rbeg + chunk_rows);
    const int lr = tid >> 4, lc0 = (tid & 15) * 8;
    float ry[8], ra[8];
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto load_regs = [&](int64_t r0) {
        const int64_t r = r0 + lr;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + lc0 + j, k = k0 + lc0 + j;
            float yv = 0.f, av = 0.f;
            if (r < rend) {
                if (n < w.N) {
                    yv = w.dy[r * w.lddy + n];
                    if (w.y_pre) yv *= act_grad(w.y_act, w.y_pre[r * w.lddy + n]);
                }
                if (k < w.K) av = fetch_a(w.a, r, k);
            }
            ry[j] = yv; ra[j] = av;
        }
    };
    auto store_regs = [&](int buf) {
        *reinterpret_cast<float4*>(&Ys[buf][lr][lc0]) = make_float4(ry[0], ry[1], ry[2], ry[3]);
        *reinterpret_cast<float4*>(&Ys[buf][lr][lc0 + 4]) = make_float4(ry[4], ry[5], ry[6], ry[7]);
        *reinterpret_cast<float4*>(&As[buf][lr][lc0]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
        *reinterpret_cast<float4*>(&As[buf][lr][lc0 + 4]) = make_float4(ra[4], ra[5], ra[6], ra[7]);
    };
    const int nsteps = (int)ceil_div<int64_t>(max((int64_t)0, rend - rbeg), 16);
    if (nsteps > 0) {
        load_regs(rbeg);
        store_regs(0);
        __syncthreads();
        for (int st = 0; st < nsteps; ++st) {
            const int buf = st & 1;
            if (st + 1 < nsteps) load_regs(rbeg + (int64_t)(st + 1) * 16);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float4 y0 = *reinterpret_cast<const float4*>(&Ys[buf][r][tn * 4]);
                const float4 y1 = *reinterpret_cast<const float4*>(&Ys[buf][r][64 + tn * 4]);
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][r][tk * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][r][64 + tk * 4]);
                const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(yv[i], av[j], acc[i][j]);
            }
            if (st + 1 < nsteps) {
                store_regs(buf ^ 1);
                __syncthreads();
            }
        }
    }
    float* out = partial + (int64_t)blockIdx.z * w.N * w.K;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + (i < 4 ? tn * 4 + i : 64 + tn * 4 + (i - 4));
        if (n >= w.N) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + (j < 4 ? tk * 4 + j : 64 + tk * 4 + (j - 4));
            if (k < w.K) out[(int64_t)n * w.K + k] = acc[i][j];
        }
    }
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int nchunks, int64_t count, int cols, float* __restrict__ out,
                       int ld_out, int accumulate) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float sum = 0.f;
    for (int c = 0; c < nchunks; ++c) sum += partial[(int64_t)c * count + i];
    int64_t r = i / cols, k = i - r * cols;
    float* o = out + r * ld_out + k;
    *o = accumulate ? *o + sum : sum;
}

static void wgrad_split(int rows, int N, int K, int* nchunks, int* chunk_rows) {
    int tiles = ceil_div(N, 128) * ceil_div(K, 128);
    int want = ceil_div(2 * sm_count(), tiles);
    int maxc = ceil_div(rows > 0 ? rows : 1, 64);
    int nc = want < 1 ? 1 : (want > maxc ? maxc : want);
    int cr = ceil_div(ceil_div(rows > 0 ? rows : 1, nc), 16) * 16;
    *chunk_rows = cr;
    *nchunks = ceil_div(rows > 0 ? rows : 1, cr);
}

size_t wgrad_workspace_bytes(int rows, int N, int K) {
    int nc, cr;
    wgrad_split(rows, N, K, &nc, &cr);
    return align_up((size_t)nc * N * K * sizeof(float)) + colsum_workspace_bytes(rows, N) + 512;
}

int launch_wgrad(const WgradArgs& w, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    int ksum = 0;
    for (int i = 0; i < w.a.nseg; ++i) ksum += w.a.k[i];
    MGB_REQUIRE(w.a.nseg >= 1 && w.a.nseg <= 4 && ksum == w.K, "wgrad: A segments (%d) do not add up to K=%d", ksum, w.K);
    int nc, cr;
    wgrad_split(w.rows, w.N, w.K, &nc, &cr);
    Workspace ws(ws_ptr, ws_bytes);
    float* partial = ws.take<float>((size_t)nc * w.N * w.K);
    MGB_WS_CHECK(ws);
    WgradArgs a = w;
    for (int i = w.a.nseg; i < 4; ++i) { a.a.p[i] = nullptr; a.a.ld[i] = 0; a.a.k[i] = 0; }
    dim3 grid(ceil_div(w.K, 128), ceil_div(w.N, 128), nc);
    wgrad_kernel<<<grid, 256, 0, s>>>(a, partial, cr);
    MGB_LAUNCH_CHECK();
    int64_t count = (int64_t)w.N * w.K;
    reduce_partials_kernel<<<(unsigned)ceil_div<int64_t>(count, 256), 256, 0, s>>>(partial, nc, count, w.K, w.dw, w.lddw, w.accumulate);
    MGB_LAUNCH_CHECK();
    if (w.db)
        MGB_TRY(launch_colsum(w.dy, w.lddy, w.y_pre, w.y_act, w.rows, w.N, w.db, w.accumulate, ws.base + ws.off,
                              ws.cap - ws.off, s));
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------
// column sums (two stages, fixed order)
// ------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 256;   // rows per block

__global__ void __launch_bounds__(128)
colsum_kernel(const float* __restrict__ x, int ld, const float* __restrict__ pre, int act, int rows, int cols,
              float* __restrict__ partial) {
    const int r0 = blockIdx.x * CS_ROWS, r1 = min(rows, r0 + CS_ROWS);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        float sum = 0.f;
        for (int r = r0; r < r1; ++r) {
            float v = x[(int64_t)r * ld + c];
            if (pre) v *= act_grad(act, pre[(int64_t)r * ld + c]);
            sum += v;
        }
        partial[(int64_t)blockIdx.x * cols + c] = sum;
    }
}

size_t colsum_workspace_bytes(int rows, int cols) {
    return align_up((size_t)ceil_div(rows > 0 ? rows : 1, CS_ROWS) * cols * sizeof(float)) + 256;
}

int launch_colsum(const float* x, int ld, const float* pre, int act, int rows, int cols, float* out, int accumulate,
                  void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    int nb = ceil_div(rows > 0 ? rows : 1, CS_ROWS);
    Workspace ws(ws_ptr, ws_bytes);
    float* partial = ws.take<float>((size_t)nb * cols);
    MGB_WS_CHECK(ws);
    colsum_kernel<<<nb, 128, 0, s>>>(x, ld, pre, act, rows, cols, partial);
    MGB_LAUNCH_CHECK();
    reduce_partials_kernel<<<ceil_div(cols, 256), 256, 0, s>>>(partial, nb, cols, cols, out, cols, accumulate);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, int ld_in, float* __restrict__ out, int ld_out) {
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
    for (int i = 0; i < 32; i += 8)
        if (r + i < rows && c < cols) tile[threadIdx.y + i][threadIdx.x] = in[(int64_t)(r + i) * ld_in + c];
    __syncthreads();
    int oc = blockIdx.y * 32 + threadIdx.x, orow = blockIdx.x * 32 + threadIdx.y;
    for (int i = 0; i < 32; i += 8)
        if (orow + i < cols && oc < rows) out[(int64_t)(orow + i) * ld_out + oc] = tile[threadIdx.x][threadIdx.y + i];
}

int launch_transpose(const float* in, int rows, int cols, int ld_in, float* out, int ld_out, cudaStream_t s) {
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(in, rows, cols, ld_in, out, ld_out);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------
// LayerNorm over 128 channels: one warp per row
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ residual, float* __restrict__ y, float* __restrict__ stats, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const float4 v = *reinterpret_cast<const float4*>(x + row * 128 + lane * 4);
    const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / 128.0f);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / 128.0f);
    const float rstd = rsqrtf(var + 1e-5f);
    const float4 gm = *reinterpret_cast<const float4*>(gamma + lane * 4);
    const float4 bt = *reinterpret_cast<const float4*>(beta + lane * 4);
    float4 o;
    o.x = dx * rstd * gm.x + bt.x; o.y = dy * rstd * gm.y + bt.y;
    o.z = dz * rstd * gm.z + bt.z; o.w = dw * rstd * gm.w + bt.w;
    if (residual) {                       // x_new + x of InteractionNetwork.forward (models/magnet_gnn.py:88) in the same pass
        const float4 r = *reinterpret_cast<const float4*>(residual + row * 128 + lane * 4);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    *reinterpret_cast<float4*>(y + row * 128 + lane * 4) = o;
    if (stats && lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
}

int launch_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, int64_t rows,
                         int cols, cudaStream_t s, const float* residual) {
    MGB_REQUIRE(cols == 128, "layernorm: only 128 channels are supported (got %d)", cols);
    if (rows == 0) return MGB_OK;
    layernorm_fwd_kernel<<<(unsigned)ceil_div<int64_t>(rows * 32, 256), 256, 0, s>>>(x, gamma, beta, residual, y, stats, rows);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

constexpr int LNB_ROWS = 64;   // rows per block in the backward kernel (8 warps x 8 rows)

__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ stats, float* __restrict__ dx, float* __restrict__ partial /*[blocks][2][128]*/,
                     int64_t rows) {
    __shared__ float red[8][2][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 gm = *reinterpret_cast<const float4*>(gamma + lane * 4);
    float dg[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < LNB_ROWS / 8; ++i) {
        const int64_t row = (int64_t)blockIdx.x * LNB_ROWS + i * 8 + warp;
        if (row >= rows) break;
        const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
        const float4 xv = *reinterpret_cast<const float4*>(x + row * 128 + lane * 4);
        const float4 gv = *reinterpret_cast<const float4*>(dy + row * 128 + lane * 4);
        const float xh[4] = {(xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd};
        const float g[4] = {gv.x * gm.x, gv.y * gm.y, gv.z * gm.z, gv.w * gm.w};
        const float gy[4] = {gv.x, gv.y, gv.z, gv.w};
        const float m1 = warp_sum(g[0] + g[1] + g[2] + g[3]) * (1.0f / 128.0f);
        const float m2 = warp_sum(g[0] * xh[0] + g[1] * xh[1] + g[2] * xh[2] + g[3] * xh[3]) * (1.0f / 128.0f);
        float4 o;
        o.x = rstd * (g[0] - m1 - xh[0] * m2); o.y = rstd * (g[1] - m1 - xh[1] * m2);
        o.z = rstd * (g[2] - m1 - xh[2] * m2); o.w = rstd * (g[3] - m1 - xh[3] * m2);
        *reinterpret_cast<float4*>(dx + row * 128 + lane * 4) = o;
#pragma unroll
        for (int j = 0; j < 4; ++j) { dg[j] += gy[j] * xh[j]; db[j] += gy[j]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[warp][0][lane * 4 + j] = dg[j]; red[warp][1][lane * 4 + j] = db[j]; }
    __syncthreads();
    const int which = threadIdx.x >> 7, c = threadIdx.x & 127;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][which][c];
    partial[((int64_t)blockIdx.x * 2 + which) * 128 + c] = sum;
}

constexpr int LNB_STAGE = 256;       // first-stage groups of the parameter-gradient reduction
size_t layernorm_bwd_workspace_bytes(int64_t rows, int cols) {
    return align_up((size_t)ceil_div<int64_t>(rows > 0 ? rows : 1, LNB_ROWS) * 2 * 128 * sizeof(float)) +
           align_up((size_t)LNB_STAGE * 2 * 128 * sizeof(float)) + 256;
}

// first stage over millions of edge rows: group g sums the partials of its contiguous range of blocks (fixed order); the
// single-block kernel below then sums the LNB_STAGE group results — one block walking 131 k partial rows took 0.86 ms
__global__ void __launch_bounds__(256)
layernorm_param_stage_kernel(const float* __restrict__ partial, int64_t nblocks, float* __restrict__ staged) {
    const int which = threadIdx.x >> 7, c = threadIdx.x & 127;
    const int64_t per = ceil_div<int64_t>(nblocks, gridDim.x);
    const int64_t b0 = (int64_t)blockIdx.x * per, b1 = b0 + per < nblocks ? b0 + per : nblocks;
    float sum = 0.f;
    for (int64_t b = b0; b < b1; ++b) sum += partial[(b * 2 + which) * 128 + c];
    staged[((int64_t)blockIdx.x * 2 + which) * 128 + c] = sum;
}

__global__ void __launch_bounds__(256)
layernorm_param_reduce_kernel(const float* __restrict__ partial, int64_t nblocks, float* __restrict__ dgamma,
                              float* __restrict__ dbeta, int accumulate) {
    const int which = threadIdx.x >> 7, c = threadIdx.x & 127;
    float sum = 0.f;
    for (int64_t b = 0; b < nblocks; ++b) sum += partial[(b * 2 + which) * 128 + c];
    float* o = (which == 0 ? dgamma : dbeta) + c;
    *o = accumulate ? *o + sum : sum;
}

int launch_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* stats, float* dx,
                         float* dgamma, float* dbeta, int accumulate_params, int64_t rows, int cols, void* ws_ptr,
                         size_t ws_bytes, cudaStream_t s) {
    MGB_REQUIRE(cols == 128, "layernorm: only 128 channels are supported (got %d)", cols);
    const int64_t nb = ceil_div<int64_t>(rows > 0 ? rows : 1, LNB_ROWS);
    Workspace ws(ws_ptr, ws_bytes);
    float* partial = ws.take<float>((size_t)nb * 2 * 128);
    float* staged = ws.take<float>((size_t)LNB_STAGE * 2 * 128);
    MGB_WS_CHECK(ws);
    if (rows == 0) { MGB_CUDA(cudaMemsetAsync(partial, 0, (size_t)nb * 2 * 128 * sizeof(float), s)); }
    else {
        layernorm_bwd_kernel<<<(unsigned)nb, 256, 0, s>>>(dy, x, gamma, stats, dx, partial, rows);
        MGB_LAUNCH_CHECK();
    }
    if (nb > 4 * LNB_STAGE) {
        layernorm_param_stage_kernel<<<LNB_STAGE, 256, 0, s>>>(partial, nb, staged);
        MGB_LAUNCH_CHECK();
        layernorm_param_reduce_kernel<<<1, 256, 0, s>>>(staged, LNB_STAGE, dgamma, dbeta, accumulate_params);
    } else {
        layernorm_param_reduce_kernel<<<1, 256, 0, s>>>(partial, nb, dgamma, dbeta, accumulate_params);
    }
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
