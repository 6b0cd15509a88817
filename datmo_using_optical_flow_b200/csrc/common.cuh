// Shared host-side plumbing of libdatmo_b200: the handle, its workspace arena,
// launch bookkeeping and per-tag event timing.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/datmo_b200.h"

struct datmo_chain;

struct datmo_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    // grow-only workspace
    char* ws = nullptr;
    size_t ws_cap = 0;
    // small pinned staging area for host-returning calls
    char* pinned = nullptr;
    size_t pinned_cap = 0;
    // grow-only device staging of the "_host" entry points (inputs / outputs of a call; the workspace
    // arena above belongs to the kernels)
    char* io = nullptr;
    size_t io_cap = 0;
    // the chain datmo_flow_to_clusters_host keeps between calls with the same configuration
    datmo_chain* chain_cache = nullptr;
    std::string err;
    // Farneback per-layer tables (pyramid taps, upsampling taps): cached on the device across calls
    // with the same geometry, so a steady-state call uploads nothing and never blocks the host
    float* fb_tab = nullptr;
    size_t fb_tab_cap = 0;          // floats
    std::vector<float> fb_tab_host;  // what fb_tab holds
    // profiling
    bool prof = false;
    unsigned prof_mask = 0xffffffffu;  // tags that get event pairs while prof is on
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    std::vector<std::pair<int, int>> ev_used;  // (tag, pool index)
    int64_t prof_launches[DATMO_TAG_COUNT] = {0};
    double prof_ms[DATMO_TAG_COUNT] = {0};
    int64_t launches = 0;
};

#define DATMO_CHECK_CUDA(h, expr)                                                              \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            char _buf[512];                                                                    \
            snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                     cudaGetErrorString(_e));                                                  \
            (h)->err = _buf;                                                                   \
            return DATMO_E_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define DATMO_REQUIRE(h, cond, msg)                  \
    do {                                             \
        if (!(cond)) {                               \
            (h)->err = std::string("invalid: ") + (msg); \
            return DATMO_E_INVALID;                  \
        }                                            \
    } while (0)

#define DATMO_ENTER(h)                                   \
    do {                                                 \
        if (!(h)) return DATMO_E_INVALID;                \
        DATMO_CHECK_CUDA(h, cudaSetDevice((h)->device)); \
    } while (0)

#define DATMO_TRY(expr)              \
    do {                             \
        int _s = (expr);             \
        if (_s != DATMO_OK) return _s; \
    } while (0)

// Bump sub-allocator over the handle's workspace.  Run once with base == nullptr to
// size the arena, then again over the real buffer.
struct Bump {
    char* base;
    size_t off = 0;
    explicit Bump(char* b) : base(b) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

int datmo_ws_reserve(datmo_ctx* h, size_t bytes);
int datmo_pinned_reserve(datmo_ctx* h, size_t bytes);
int datmo_io_reserve(datmo_ctx* h, size_t bytes);

// RAII-ish bracket around a kernel launch: counts it and, when profiling is on,
// records an event pair on the handle's stream.
struct LaunchScope {
    datmo_ctx* h;
    int idx = -1;
    LaunchScope(datmo_ctx* h_, int tag, int n_launches = 1);
    ~LaunchScope();
};

// Opt-in dynamic shared memory is a per-device function attribute: remember, per call site and per
// device, the largest size already granted (a process may drive several GPUs through several handles).
constexpr int DATMO_MAX_DEVICES = 64;
struct SmemGrant {
    size_t granted[DATMO_MAX_DEVICES] = {0};
};
template <typename Kernel>
int datmo_grant_smem(datmo_ctx* h, Kernel kernel, size_t bytes, SmemGrant& g) {
    if (bytes <= 48 * 1024) return DATMO_OK;   // no opt-in needed
    const int d = h->device >= 0 && h->device < DATMO_MAX_DEVICES ? h->device : 0;
    if (bytes <= g.granted[d] && h->device < DATMO_MAX_DEVICES) return DATMO_OK;
    DATMO_CHECK_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    g.granted[d] = bytes;
    return DATMO_OK;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// dbscan.cu: exclusive rank of every flagged item of [batch][n]; block_sums needs batch * ceil(n / 4096) ints
int datmo_flag_scan(datmo_ctx* h, const uint8_t* flags, int64_t n, int batch, int32_t* block_sums, int32_t* totals,
                    int32_t* rank, int tag, int sparse);

// dbscan_runs.cu: the run-based DBSCAN (1 <= floor(eps) <= 15); called by datmo_dbscan_grid_dev
bool datmo_dbscan_runs_supported(double eps);
size_t datmo_dbscan_runs_workspace(int H, int W, int batch);
int datmo_dbscan_runs(datmo_ctx* h, char* ws, const float* vx_f, const float* vy_f, const uint8_t* valid, int H, int W,
                      int batch, double eps, int min_samples, int cap, int32_t* n_valid, int32_t* labels,
                      int32_t* indices, int32_t* n_clusters, int (*tag)(int));

#ifdef __CUDACC__
// Order-independent accumulation of a partial sum p: p = hi + lo / 2^64 with hi = floor(p) (int64) and lo the
// fraction as a 64-bit integer, added with two integer atomics (the low word's wrap-around carries into
// the high one).  Integer addition commutes, so the total does not depend on the order in which warps
// and CTAs arrive — unlike an fp64 atomicAdd, whose result changes in the last bits from run to run.
// Bits of p below 2^-64 are dropped (the velocities are f32: nothing below 2^-40 of a value >= 2^-17
// exists); a non-finite or astronomically large partial goes to an fp64 accumulator instead.
__device__ __forceinline__ void datmo_fixed_add(unsigned long long* hi_acc, unsigned long long* lo_acc,
                                          unsigned long long* overflow_acc, double p) {
    if (!(fabs(p) < 4.0e18)) {
        atomicAdd(reinterpret_cast<double*>(overflow_acc), p);
        return;
    }
    const double fl = floor(p);
    const long long hi = static_cast<long long>(fl);
    const unsigned long long lo = __double2ull_rz((p - fl) * 18446744073709551616.0);
    const unsigned long long old = atomicAdd(lo_acc, lo);
    const unsigned long long carry = old + lo < old ? 1ull : 0ull;
    atomicAdd(hi_acc, static_cast<unsigned long long>(hi) + carry);
}

// value of such an accumulator: (signed high word) + (low word) / 2^64 + the fp64 side accumulator
__device__ __forceinline__ double datmo_fixed_value(unsigned long long hi, unsigned long long lo, double overflow) {
    return static_cast<double>(static_cast<long long>(hi)) + static_cast<double>(lo) * 5.421010862427522e-20 + overflow;
}

#endif

#define DATMO_POST_LAUNCH(h) DATMO_CHECK_CUDA(h, cudaGetLastError())
