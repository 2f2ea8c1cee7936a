// magnet_b200 — shared helpers for the sm_100a kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define MGB_OK 0
#define MGB_ERR_ARG (-1)        // bad argument
#define MGB_ERR_WORKSPACE (-2)  // workspace too small
#define MGB_ERR_CAPACITY (-3)   // output capacity overflow
#define MGB_ERR_CUDA (-4)       // CUDA runtime error

namespace mgb {

void set_error(const char* fmt, ...);   // thread-local message, read through mgb_last_error()

#define MGB_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            mgb::set_error("%s:%d CUDA error %s (%s)", __FILE__, __LINE__, cudaGetErrorName(_e), \
                           cudaGetErrorString(_e));                                          \
            return MGB_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

// every kernel launch is followed by exactly one MGB_LAUNCH_CHECK(): it also feeds the launch counter
// that bench.py reports as "gpu_launches"
void count_launch();
long long launch_count();
#define MGB_LAUNCH_CHECK()          \
    do {                            \
        mgb::count_launch();        \
        MGB_CUDA(cudaGetLastError()); \
    } while (0)

#define MGB_REQUIRE(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            mgb::set_error(__VA_ARGS__);      \
            return MGB_ERR_ARG;               \
        }                                     \
    } while (0)

#define MGB_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != MGB_OK) return _rc; \
    } while (0)

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Bump allocator over the caller-owned workspace (the library never allocates device memory).
struct Workspace {
    char* base;
    size_t cap;
    size_t off;
    bool ok;
    Workspace(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), ok(true) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = align_up(n * sizeof(T));
        if (base == nullptr || off + bytes > cap) { ok = false; off += bytes; return nullptr; }
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
};

#define MGB_WS_CHECK(ws)                                                                   \
    do {                                                                                   \
        if (!(ws).ok) {                                                                    \
            mgb::set_error("%s:%d workspace too small: need %zu bytes, have %zu", __FILE__, \
                           __LINE__, (ws).off, (ws).cap);                                  \
            return MGB_ERR_WORKSPACE;                                                      \
        }                                                                                  \
    } while (0)

// ---- optional per-kernel timing (bench.py roofline): CUDA events on the launching stream -------
enum ProfId : int { PROF_EDGE_FWD = 0, PROF_EDGE_BWD = 1, PROF_NODE_GEMM = 2, PROF_WGRAD = 3, PROF_GRAPH = 4,
                    PROF_IN_EDGE_FWD = 5, PROF_IN_EDGE_BWD = 6, PROF_INR_DECODE = 7, PROF_COUNT = 8 };
void prof_begin(int id, cudaStream_t s);
void prof_end(int id, cudaStream_t s);
struct ProfScope {
    int id; cudaStream_t s;
    ProfScope(int id_, cudaStream_t s_) : id(id_), s(s_) { prof_begin(id, s); }
    ~ProfScope() { prof_end(id, s); }
};

int sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device

// ---- scan / sort primitives (scan_sort.cu) ------------------------------------------------
size_t scan_workspace_bytes(int64_t n);
// out[i] = sum_{j<i} in[j], out has n+1 entries (out[n] = total). in/out may not alias.
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes, cudaStream_t s);
size_t sort_workspace_bytes(int64_t n);
// Stable LSD radix sort of (key, val) pairs on the low `bits` bits of key. Result in keys_out/vals_out.
int radix_sort_pairs(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                     int64_t n, int bits, void* ws, size_t ws_bytes, cudaStream_t s);
// starts[c] = first position p with sorted_keys[p] >= c, for c in [0, n_keys]; starts[n_keys] = n.
int segment_starts(const uint32_t* sorted_keys, int64_t n, int32_t* starts, int64_t n_keys, cudaStream_t s);

}  // namespace mgb
