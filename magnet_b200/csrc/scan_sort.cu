// magnet_b200 — device-wide exclusive scan and stable LSD radix sort (hand-written; no CUB).
//
// These are the integer building blocks of the graph builder (cell binning, CSR plans).  They
// are HBM-bound streaming passes: coalesced 128-bit loads, shared-memory staging for the
// ranks, grids sized by the tile count.  Everything is deterministic (no float atomics; the
// only atomics are integer histogram counters whose result is order independent).
#include "common.cuh"
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>

namespace mgb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ---- event profiler -----------------------------------------------------------------------------
static std::mutex g_prof_mu;
static bool g_prof_on = false;
struct ProfRec { cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof[PROF_COUNT];
static cudaEvent_t g_prof_open[PROF_COUNT];
void prof_enable(bool on) { std::lock_guard<std::mutex> l(g_prof_mu); g_prof_on = on; }
void prof_begin(int id, cudaStream_t s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> l(g_prof_mu);
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    g_prof_open[id] = e;
}
void prof_end(int id, cudaStream_t s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> l(g_prof_mu);
    cudaEvent_t e;
    if (g_prof_open[id] == nullptr || cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    g_prof[id].push_back({g_prof_open[id], e});
    g_prof_open[id] = nullptr;
}
// synchronises the recorded events, returns total milliseconds and launch count, and clears the slot
int prof_collect(int id, double* total_ms, long long* count) {
    std::lock_guard<std::mutex> l(g_prof_mu);
    double t = 0;
    long long c = 0;
    for (auto& r : g_prof[id]) {
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { t += ms; ++c; }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof[id].clear();
    *total_ms = t;
    *count = c;
    return MGB_OK;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ------------------------------------------------------------------------------------------
// exclusive scan (int32), 2048 items per block, recursive over block sums
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* smem_warp /*[8]*/, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    int wsum = (lane < 8) ? smem_warp[lane] : 0;
    int winc = wsum;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
    }
    int wbase = __shfl_sync(0xffffffffu, winc - wsum, warp);
    *total = __shfl_sync(0xffffffffu, winc, 7);
    __syncthreads();
    return wbase + inc - v;
}

// processes positions [0, n]: in[n] is read as 0 so that out[n] = total.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t* __restrict__ tile_sums,
                 int64_t n) {
    __shared__ int warp_sums[8];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t p = base + i;
        v[i] = (p < n) ? in[p] : 0;
        sum += v[i];
    }
    int total;
    int excl = block_exclusive_scan_256(sum, warp_sums, &total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t p = base + i;
        if (p <= n) out[p] = excl;
        excl += v[i];
    }
    if (threadIdx.x == 0 && tile_sums) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ tile_offsets, int64_t n) {
    const int off = tile_offsets[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t p = base + i;
        if (p <= n) out[p] += off;
    }
}

size_t scan_workspace_bytes(int64_t n) {
    size_t bytes = 0;
    int64_t m = n + 1;
    while (m > SCAN_TILE) {
        int64_t tiles = ceil_div<int64_t>(m, SCAN_TILE);
        bytes += 2 * align_up((size_t)(tiles + 1) * sizeof(int32_t));
        m = tiles + 1;
    }
    return bytes + 256;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    if (n < 0) { set_error("scan: negative length"); return MGB_ERR_ARG; }
    const int64_t m = n + 1;
    const int64_t tiles = ceil_div<int64_t>(m, SCAN_TILE);
    if (tiles == 1) {
        scan_tile_kernel<<<1, SCAN_THREADS, 0, s>>>(in, out, nullptr, n);
        MGB_LAUNCH_CHECK();
        return MGB_OK;
    }
    Workspace ws(ws_ptr, ws_bytes);
    int32_t* sums = ws.take<int32_t>(tiles + 1);
    int32_t* offs = ws.take<int32_t>(tiles + 1);
    MGB_WS_CHECK(ws);
    scan_tile_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, out, sums, n);
    MGB_LAUNCH_CHECK();
    MGB_TRY(exclusive_scan_i32(sums, offs, tiles, ws.base + ws.off, ws.cap - ws.off, s));
    scan_add_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(out, offs, n);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass, 2048 items per block
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;                       // rounds per warp
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 2048
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int32_t* __restrict__ hist /*[256][tiles]*/,
                  int64_t tiles) {
    __shared__ int cnt[RS_RADIX];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        int64_t p = base + (int64_t)i * RS_THREADS + threadIdx.x;
        if (p < n) atomicAdd(&cnt[(keys[p] >> shift) & 0xff], 1);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * tiles + blockIdx.x] = cnt[threadIdx.x];
}

// Item order inside a tile: warp w owns items [w*256, w*256+256); round r covers [r*32, r*32+32).
__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                     const int32_t* __restrict__ hist_scan /*[256][tiles]*/, int64_t tiles) {
    __shared__ int warp_cnt[8][RS_RADIX];   // running count of each digit inside each warp
    __shared__ int digit_base[RS_RADIX];    // global destination of the tile's first item of each digit
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * RS_RADIX; i += RS_THREADS) (&warp_cnt[0][0])[i] = 0;
    digit_base[threadIdx.x] = hist_scan[(int64_t)threadIdx.x * tiles + blockIdx.x];
    __syncthreads();

    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (RS_ITEMS * 32);
    uint32_t key[RS_ITEMS], val[RS_ITEMS];
    int rank[RS_ITEMS];
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t p = wbase + r * 32 + lane;
        bool valid = p < n;
        key[r] = valid ? keys_in[p] : 0xffffffffu;
        val[r] = valid ? vals_in[p] : 0u;
        int d = valid ? (int)((key[r] >> shift) & 0xff) : 256 + lane;   // invalid lanes never match anyone
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int prev = valid ? warp_cnt[warp][d] : 0;
        rank[r] = prev + __popc(peers & lt_mask);
        __syncwarp();
        if (valid && (peers & lt_mask) == 0) warp_cnt[warp][d] = prev + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over the 8 warps for digit = threadIdx.x
    {
        int run = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            int c = warp_cnt[w][threadIdx.x];
            warp_cnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t p = wbase + r * 32 + lane;
        if (p < n) {
            int d = (int)((key[r] >> shift) & 0xff);
            int64_t dst = (int64_t)digit_base[d] + warp_cnt[warp][d] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = val[r];
        }
    }
}

size_t sort_workspace_bytes(int64_t n) {
    int64_t tiles = ceil_div<int64_t>(n > 0 ? n : 1, RS_TILE);
    int64_t hn = tiles * RS_RADIX;
    return 2 * align_up((size_t)(hn + 1) * 4) + 2 * align_up((size_t)(n > 0 ? n : 1) * 4) + scan_workspace_bytes(hn) + 1024;
}

int radix_sort_pairs(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                     int64_t n, int bits, void* ws_ptr, size_t ws_bytes, cudaStream_t s) {
    if (n == 0) return MGB_OK;
    if (n >= (int64_t)1 << 31) { set_error("sort: more than 2^31-1 items"); return MGB_ERR_ARG; }
    if (bits < 1) bits = 1;
    if (bits > 32) bits = 32;
    const int passes = ceil_div(bits, 8);
    const int64_t tiles = ceil_div<int64_t>(n, RS_TILE);
    const int64_t hn = tiles * RS_RADIX;
    Workspace ws(ws_ptr, ws_bytes);
    int32_t* hist = ws.take<int32_t>(hn + 1);
    int32_t* hscan = ws.take<int32_t>(hn + 1);
    uint32_t* tk = ws.take<uint32_t>(n);
    uint32_t* tv = ws.take<uint32_t>(n);
    MGB_WS_CHECK(ws);
    void* scan_ws = ws.base + ws.off;
    size_t scan_ws_bytes = ws.cap - ws.off;
    // ping-pong so that the last pass lands in keys_out/vals_out
    const uint32_t* src_k = keys_in;
    const uint32_t* src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        bool to_out = ((passes - 1 - p) % 2) == 0;
        uint32_t* dst_k = to_out ? keys_out : tk;
        uint32_t* dst_v = to_out ? vals_out : tv;
        radix_hist_kernel<<<(unsigned)tiles, RS_THREADS, 0, s>>>(src_k, n, 8 * p, hist, tiles);
        MGB_LAUNCH_CHECK();
        MGB_TRY(exclusive_scan_i32(hist, hscan, hn, scan_ws, scan_ws_bytes, s));
        radix_scatter_kernel<<<(unsigned)tiles, RS_THREADS, 0, s>>>(src_k, src_v, dst_k, dst_v, n, 8 * p, hscan, tiles);
        MGB_LAUNCH_CHECK();
        src_k = dst_k;
        src_v = dst_v;
    }
    return MGB_OK;
}

__global__ void segment_starts_kernel(const uint32_t* __restrict__ keys, int64_t n, int32_t* __restrict__ starts,
                                      int64_t n_keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 || i >= n) return;
    // position i closes the key range (prev, cur]: every key c in it starts at i (gaps between two present keys are short)
    const int64_t prev = (int64_t)keys[i - 1];
    int64_t cur = (int64_t)keys[i];
    if (cur > n_keys) cur = n_keys;
    for (int64_t c = prev + 1; c <= cur; ++c) starts[c] = (int32_t)i;
}

// the two open ends — keys up to the first present key start at 0, keys past the last present key start at n — can
// span most of a sparse key space (a kNN grid with millions of cells): one thread per key instead of one thread's loop
__global__ void segment_ends_kernel(const uint32_t* __restrict__ keys, int64_t n, int32_t* __restrict__ starts, int64_t n_keys) {
    const int64_t first = n > 0 ? (int64_t)keys[0] : n_keys;
    const int64_t last = n > 0 ? (int64_t)keys[n - 1] : -1;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= n_keys; c += (int64_t)gridDim.x * blockDim.x) {
        if (c <= first) starts[c] = 0;
        else if (c > last) starts[c] = (int32_t)n;
    }
}

int segment_starts(const uint32_t* sorted_keys, int64_t n, int32_t* starts, int64_t n_keys, cudaStream_t s) {
    if (n > 1) {
        segment_starts_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, s>>>(sorted_keys, n, starts, n_keys);
        MGB_LAUNCH_CHECK();
    }
    const int64_t blocks = ceil_div<int64_t>(n_keys + 1, 256);
    segment_ends_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, s>>>(sorted_keys, n, starts, n_keys);
    MGB_LAUNCH_CHECK();
    return MGB_OK;
}

}  // namespace mgb
