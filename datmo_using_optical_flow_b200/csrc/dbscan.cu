// DBSCAN over the (row, col, vx, vy) features of the valid BEV cells, as a grid
// union-find, plus the per-cluster summaries.  sm_100a.
//
// Replaces dbscan_clustering (Optical_flow/main.py:231-259: sklearn.cluster.DBSCAN on
// the row-major list of valid cells) and extract_cluster_data (main.py:402-434).
// sklearn's result is reproduced exactly, numbering included, by the grid rule of
// SURVEY.md §8 a8 (restated in oracle/dbscan_np.py):
//   neighbours    valid cells of the (2*floor(eps)+1)^2 window with
//                 d2 = drow^2 + dcol^2 + dvx^2 + dvy^2 <= eps^2 in fp64, that summation
//                 order, every operation rounded (no FMA); self included
//   core          neighbour count >= min_samples
//   clusters      connected components of core cells; root = MINIMUM row-major index
//   border cell   takes the smallest root among its core neighbours, else noise (-1)
//   label         rank of the root among all roots in ascending order
// Row-major ranks (the order of np.nonzero) come from a flag scan of the grid.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "dbscan_common.cuh"

namespace {

constexpr int SCAN_ITEMS = 4096;  // grid cells handled by one CTA of the flag scan
constexpr int SCAN_THREADS = 256;

// ---- generic exclusive scan of uint8 flags over [batch][n] ---------------------------
// Each thread owns 16 consecutive flags, fetched with one 16-byte load when the frame is
// 16-byte aligned (n % 16 == 0 and an aligned base); bit k of the result = flag k is non-zero.
__device__ __forceinline__ unsigned load_flags16(const uint8_t* __restrict__ f, int64_t p0, int64_t n, bool vec) {
    unsigned bits = 0;
    if (vec && p0 + 16 <= n) {
        const uint4 v = *reinterpret_cast<const uint4*>(f + p0);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned m = __vcmpne4(w[j], 0u);  // 0xff per non-zero byte
            bits |= ((m & 1u) | ((m >> 7) & 2u) | ((m >> 14) & 4u) | ((m >> 21) & 8u)) << (4 * j);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (p0 + i < n && f[p0 + i] != 0) bits |= 1u << i;
    }
    return bits;
}

// pass 1: per-CTA totals
__global__ void __launch_bounds__(SCAN_THREADS) k_flag_block_sums(const uint8_t* __restrict__ flags, int64_t n,
                                                                  int nblk, int32_t* __restrict__ block_sums,
                                                                  int vec) {
    const int blk = blockIdx.x, b = blockIdx.y;
    const uint8_t* f = flags + static_cast<size_t>(b) * n;
    static_assert(SCAN_ITEMS == 16 * SCAN_THREADS, "one 16-flag group per thread");
    const int64_t p0 = static_cast<int64_t>(blk) * SCAN_ITEMS + static_cast<int64_t>(threadIdx.x) * 16;
    int cnt = p0 < n ? __popc(load_flags16(f, p0, n, vec != 0)) : 0;
    __shared__ int s_w[SCAN_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < SCAN_THREADS / 32; ++i) t += s_w[i];
        block_sums[static_cast<size_t>(b) * nblk + blk] = t;
    }
}

// pass 2: one CTA per frame turns the CTA totals into exclusive offsets and the total
__global__ void __launch_bounds__(256) k_scan_block_sums(int32_t* __restrict__ block_sums, int nblk,
                                                         int32_t* __restrict__ totals) {
    const int b = blockIdx.x;
    int32_t* s = block_sums + static_cast<size_t>(b) * nblk;
    __shared__ int s_part[256];
    // each thread owns a contiguous run
    int per = (nblk + 255) / 256;
    int lo = threadIdx.x * per, hi = min(lo + per, nblk);
    int t = 0;
    for (int i = lo; i < hi; ++i) t += s[i];
    s_part[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < 256; ++i) {
            int v = s_part[i];
            s_part[i] = run;
            run += v;
        }
        totals[b] = run;
    }
    __syncthreads();
    int run = s_part[threadIdx.x];
    for (int i = lo; i < hi; ++i) {
        int v = s[i];
        s[i] = run;
        run += v;
    }
}

// pass 3: exclusive rank of every flagged cell (others get -1)
__global__ void __launch_bounds__(SCAN_THREADS) k_flag_ranks(const uint8_t* __restrict__ flags, int64_t n, int nblk,
                                                             const int32_t* __restrict__ block_offs,
                                                             int32_t* __restrict__ rank, int sparse, int vec) {
    const int blk = blockIdx.x, b = blockIdx.y;
    const uint8_t* f = flags + static_cast<size_t>(b) * n;
    int32_t* r = rank + static_cast<size_t>(b) * n;
    constexpr int PER = SCAN_ITEMS / SCAN_THREADS;  // contiguous cells per thread
    static_assert(PER == 16, "one 16-flag group per thread");
    const int64_t p0 = static_cast<int64_t>(blk) * SCAN_ITEMS + static_cast<int64_t>(threadIdx.x) * PER;
    const unsigned bits = p0 < n ? load_flags16(f, p0, n, vec != 0) : 0u;
    const int cnt = __popc(bits);
    // CTA-wide exclusive scan of cnt
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __shared__ int s_w[SCAN_THREADS / 32];
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    int woff = 0;
    for (int i = 0; i < wid; ++i) woff += s_w[i];
    int run = block_offs[static_cast<size_t>(b) * nblk + blk] + woff + inc - cnt;
    if (sparse) {
        // only flagged cells are ever looked up: skip the 4 B/cell fill
        unsigned rem = bits;
        while (rem) {
            const int i = __ffs(rem) - 1;
            rem &= rem - 1;
            r[p0 + i] = run++;
        }
    } else {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            if (p0 + i < n) {
                const int on = (bits >> i) & 1;
                r[p0 + i] = on ? run : -1;
                run += on;
            }
        }
    }
}

// state: 0 = not valid, 1 = valid non-core, 2 = core.  parent: self for core cells, -1 otherwise.
__device__ __forceinline__ void core_cell(int x, int y, int b, const float* __restrict__ vx, const float* __restrict__ vy,
                                              const uint8_t* __restrict__ valid, int H, int W, int r, EpsTest eps2,
                                              int min_samples, uint8_t* __restrict__ state,
                                              int32_t* __restrict__ parent) {

    const size_t base = static_cast<size_t>(b) * H * W;
    const int o = y * W + x;
    uint8_t st = 0;
    if (valid[base + o]) {
        const float vx0 = vx[base + o], vy0 = vy[base + o];
        // nearest neighbours first: inside a moving region self + the 4-neighbourhood usually
        // reaches min_samples; only otherwise count over the whole window
        int cnt = 1;  // self
        if (r >= 1 && cnt < min_samples) {
            // all four flags, then all four velocity pairs: two load latencies instead of eight
            const int ndr[4] = {0, 0, -1, 1}, ndc[4] = {-1, 1, 0, 0};
            bool on[4];
            float nvx[4], nvy[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int yy = y + ndr[k], xx = x + ndc[k];
                on[k] = yy >= 0 && yy < H && xx >= 0 && xx < W && valid[base + static_cast<size_t>(yy) * W + xx];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const size_t q = base + static_cast<size_t>(y + ndr[k]) * W + (x + ndc[k]);
                nvx[k] = on[k] ? vx[q] : 0.f;
                nvy[k] = on[k] ? vy[q] : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (on[k] && within_eps(ndr[k], ndc[k], vx0, vy0, nvx[k], nvy[k], eps2)) ++cnt;
        }
        if (cnt < min_samples) cnt = 0;  // inconclusive: recount exactly
        for (int dr = -r; dr <= r && cnt < min_samples; ++dr) {
            int yy = y + dr;
            if (yy < 0 || yy >= H) continue;
            for (int dc = -r; dc <= r; ++dc) {
                int xx = x + dc;
                if (xx < 0 || xx >= W) continue;
                size_t q = base + static_cast<size_t>(yy) * W + xx;
                if (valid[q] && within_eps(dr, dc, vx0, vy0, vx[q], vy[q], eps2)) {
                    if (++cnt >= min_samples) break;
                }
            }
        }
        st = cnt >= min_samples ? 2 : 1;
    }
    state[base + o] = st;
    parent[base + o] = st == 2 ? o : -1;
}

// Thread-per-cell kernels of this file handle FOUR consecutive cells per thread: one vector load
// tells whether any of them has work (four in five do not), which quarters the number of CTAs the
// GPU has to turn over — with one cell per thread that turnover, not memory, was the bound.
__device__ __forceinline__ bool vec4_ok(int W, const void* p) {
    return (W & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

__global__ void __launch_bounds__(256) k_core(const float* __restrict__ vx, const float* __restrict__ vy,
                                              const uint8_t* __restrict__ valid, int H, int W, int r, EpsTest eps2,
                                              int min_samples, uint8_t* __restrict__ state,
                                              int32_t* __restrict__ parent) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x4 >= W) return;
    const size_t o4 = (static_cast<size_t>(b) * H + y) * W + x4;
    if (vec4_ok(W, valid) && vec4_ok(W, state) && vec4_ok(W, parent) &&
        *reinterpret_cast<const unsigned*>(valid + o4) == 0u) {
        *reinterpret_cast<unsigned*>(state + o4) = 0u;
        *reinterpret_cast<int4*>(parent + o4) = make_int4(-1, -1, -1, -1);
        return;
    }
    for (int j = 0; j < 4 && x4 + j < W; ++j) core_cell(x4 + j, y, b, vx, vy, valid, H, W, r, eps2, min_samples, state, parent);
}

// Link pass 1 (no atomics): every core cell points at the smallest-index core cell among its
// four preceding 8-neighbours that is within eps (or stays its own root).  All candidates
// precede the cell in row-major order, so the result is a forest, and an 8-connected region
// of mutually close cells collapses into ONE tree (diagonal chains end on its top row / left
// column, which chain to its first cell).  Linking to the first cell of the whole half-window
// instead was measured slower: its chains stride (-4,-2) and stay interleaved, leaving several
// trees per region and defeating the quick reject of pass 2.
__device__ __forceinline__ void link_near_cell(int x, int y, int b, const float* __restrict__ vx, const float* __restrict__ vy,
                                                   const uint8_t* __restrict__ state, int H, int W, int r,
                                                   EpsTest eps2, int32_t* __restrict__ parent) {

    const size_t base = static_cast<size_t>(b) * H * W;
    const int o = y * W + x;
    if (state[base + o] != 2) return;
    const float vx0 = vx[base + o], vy0 = vy[base + o];
    int best = o;
    // ascending index order: (-1,-1), (-1,0), (-1,+1), (0,-1); keep the first hit.  States first,
    // then the velocities of the core ones: independent loads, two latencies in all.
    const int ndr[4] = {-1, -1, -1, 0}, ndc[4] = {-1, 0, 1, -1};
    bool on[4];
    float nvx[4], nvy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int yy = y + ndr[k], xx = x + ndc[k];
        on[k] = yy >= 0 && xx >= 0 && xx < W && state[base + yy * W + xx] == 2;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int q = (y + ndr[k]) * W + x + ndc[k];
        nvx[k] = on[k] ? vx[base + q] : 0.f;
        nvy[k] = on[k] ? vy[base + q] : 0.f;
    }
#pragma unroll
    for (int k = 3; k >= 0; --k)
        if (on[k] && within_eps(ndr[k], ndc[k], vx0, vy0, nvx[k], nvy[k], eps2)) best = (y + ndr[k]) * W + x + ndc[k];
    parent[base + o] = best;
}

__global__ void __launch_bounds__(256) k_link_near(const float* __restrict__ vx, const float* __restrict__ vy,
                                                   const uint8_t* __restrict__ state, int H, int W, int r,
                                                   EpsTest eps2, int32_t* __restrict__ parent) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x4 >= W || r < 1) return;
    if (vec4_ok(W, state)) {
        const unsigned st4 = *reinterpret_cast<const unsigned*>(state + (static_cast<size_t>(b) * H + y) * W + x4);
        if ((st4 & 0x02020202u) == 0u) return;  // no core cell among the four
    }
    for (int j = 0; j < 4 && x4 + j < W; ++j) link_near_cell(x4 + j, y, b, vx, vy, state, H, W, r, eps2, parent);
}

// Link pass 1b (after a flatten): the atomics-free pass leaves an 8-connected region split
// into a few diagonal stripes (one per ragged top / left edge cell).  Join them: same four
// neighbours, but now "same flattened parent" skips the pair, so only cells ON a stripe
// interface reach the atomic path.  Afterwards trees == 8-connected regions of close cells.
__device__ __forceinline__ void union_near_cell(int x, int y, int b, const float* __restrict__ vx, const float* __restrict__ vy,
                                                    const uint8_t* __restrict__ state, int H, int W, int r,
                                                    EpsTest eps2, int32_t* __restrict__ parent) {

    const size_t base = static_cast<size_t>(b) * H * W;
    const int o = y * W + x;
    if (state[base + o] != 2) return;
    int32_t* par = parent + base;
    const int my_root = par[o];
    int joined = -1;
    const float vx0 = vx[base + o], vy0 = vy[base + o];
    const int ndr[4] = {-1, -1, -1, 0}, ndc[4] = {-1, 0, 1, -1};
    // the four neighbours' parents in one go; velocities only for those in a different tree
    int pq[4];
    float nvx[4], nvy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int yy = y + ndr[k], xx = x + ndc[k];
        pq[k] = (yy >= 0 && xx >= 0 && xx < W) ? par[yy * W + xx] : -1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool want = pq[k] >= 0 && pq[k] != my_root;
        const int q = (y + ndr[k]) * W + x + ndc[k];
        nvx[k] = want ? vx[base + q] : 0.f;
        nvy[k] = want ? vy[base + q] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (pq[k] < 0 || pq[k] == my_root || pq[k] == joined) continue;
        if (within_eps(ndr[k], ndc[k], vx0, vy0, nvx[k], nvy[k], eps2)) {
            uf_union(par, my_root, pq[k]);  // both are members of the two trees: start the finds one hop up
            joined = pq[k];
        }
    }
}

__global__ void __launch_bounds__(256) k_union_near(const float* __restrict__ vx, const float* __restrict__ vy,
                                                    const uint8_t* __restrict__ state, int H, int W, int r,
                                                    EpsTest eps2, int32_t* __restrict__ parent) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x4 >= W || r < 1) return;
    if (vec4_ok(W, state)) {
        const unsigned st4 = *reinterpret_cast<const unsigned*>(state + (static_cast<size_t>(b) * H + y) * W + x4);
        if ((st4 & 0x02020202u) == 0u) return;
    }
    for (int j = 0; j < 4 && x4 + j < W; ++j) union_near_cell(x4 + j, y, b, vx, vy, state, H, W, r, eps2, parent);
}

// Link pass 2 (after a flatten): the whole preceding half-window.  parent[] now holds each
// cell's pass-1 root, so "same parent" proves "same component" from one cached load and
// skips both the fp64 distance test and the union; only the few pairs that bridge different
// pass-1 trees reach the atomic path.
__global__ void __launch_bounds__(256) k_union_far(const float* __restrict__ vx, const float* __restrict__ vy,
                                                   const uint8_t* __restrict__ state, int H, int W, int r,
                                                   EpsTest eps2, int32_t* __restrict__ parent) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    const size_t base = static_cast<size_t>(b) * H * W;
    int32_t* par = parent + base;
    const int o = y * W + x;
    const bool core = x < W && state[base + o] == 2;
    if (!__any_sync(0xffffffffu, core)) return;  // three warps in four hold no core cell at all
    // Warp-level quick reject.  The 32 cells of a warp share the footprint rows y-r..y,
    // columns x0-r..x0+31+r.  If every core cell in that footprint already has the same
    // pass-1 root there is nothing to join and the whole warp leaves after ~2(r+1) loads per
    // lane instead of walking (r+1)(2r+1) neighbours each.
    {
        const int lane = threadIdx.x & 31;
        const int xw0 = x - lane;  // first column of this warp
        int lo = 0x7fffffff, hi = -1;
        constexpr int QR = 5;  // windows up to this radius (the reference's eps 5) fetch the whole footprint before looking at any of it
        if (r <= QR) {
            // 2 (r + 1) independent loads per lane: one memory latency for the footprint, not one per load
            int p0[QR + 1], p1[QR + 1];
            const int ca = xw0 - r + lane, cb = ca + 32;
            const bool oka = ca >= 0 && ca < W, okb = cb <= xw0 + 31 + r && cb < W;
#pragma unroll
            for (int t = 0; t <= QR; ++t) {
                const int yy = y - t;
                const bool row = t <= r && yy >= 0;
                p0[t] = row && oka ? par[yy * W + ca] : -1;
                p1[t] = row && okb ? par[yy * W + cb] : -1;
            }
#pragma unroll
            for (int t = 0; t <= QR; ++t) {
                if (p0[t] >= 0) lo = min(lo, p0[t]), hi = max(hi, p0[t]);
                if (p1[t] >= 0) lo = min(lo, p1[t]), hi = max(hi, p1[t]);
            }
        } else {
            for (int dr = -r; dr <= 0; ++dr) {
                const int yy = y + dr;
                if (yy < 0) continue;
                const int32_t* prow = par + yy * W;
                for (int cx = xw0 - r + lane; cx <= xw0 + 31 + r; cx += 32) {
                    if (cx < 0 || cx >= W) continue;
                    const int p = prow[cx];
                    if (p >= 0) {
                        lo = min(lo, p);
                        hi = max(hi, p);
                    }
                }
            }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (hi < 0 || lo == hi) return;
    }
    if (!core) return;
    const float vx0 = vx[base + o], vy0 = vy[base + o];
    const int my_root = par[o];
    int joined = -1;  // the last foreign pass-1 tree this cell has already been united with
    for (int dr = -r; dr <= 0; ++dr) {
        const int yy = y + dr;
        if (yy < 0) continue;
        const int dc_lo = max(-r, -x), dc_hi = dr == 0 ? -1 : min(r, W - 1 - x);
        const int32_t* prow = par + yy * W + x;
        // parent is -1 for anything that is not a core cell, so one load filters
        // "not core", "same pass-1 tree" and "tree already joined"
        constexpr int QR = 5;
        if (r <= QR) {
            // the row's candidates are fetched together (independent loads), then examined
            int pq[2 * QR + 1];
#pragma unroll
            for (int t = 0; t < 2 * QR + 1; ++t) {
                const int dc = t - QR;
                pq[t] = dc >= dc_lo && dc <= dc_hi ? prow[dc] : -1;
            }
#pragma unroll
            for (int t = 0; t < 2 * QR + 1; ++t) {
                if (pq[t] < 0 || pq[t] == my_root || pq[t] == joined) continue;
                const int dc = t - QR;
                const int q = yy * W + x + dc;
                if (within_eps(dr, dc, vx0, vy0, vx[base + q], vy[base + q], eps2)) {
                    uf_union(par, my_root, pq[t]);  // members of the two trees, one hop closer to the roots
                    joined = pq[t];
                }
            }
        } else {
            for (int dc = dc_lo; dc <= dc_hi; ++dc) {
                const int pq = prow[dc];
                if (pq < 0 || pq == my_root || pq == joined) continue;
                const int q = yy * W + x + dc;
                if (within_eps(dr, dc, vx0, vy0, vx[base + q], vy[base + q], eps2)) {
                    uf_union(par, my_root, pq);
                    joined = pq;
                }
            }
        }
    }
}

// Flatten: point every core cell straight at its root (parent is -1 for anything that is not a
// core cell, so the forest alone says who takes part).  Four cells per thread through one
// 16-byte load; MARK_ROOTS additionally writes the root flags the final label scan consumes.
template <bool MARK_ROOTS>
__global__ void __launch_bounds__(256) k_flatten(int64_t n, int32_t* __restrict__ parent,
                                                 uint8_t* __restrict__ is_root) {
    const int64_t i4 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    const int b = blockIdx.y;
    if (i4 >= n) return;
    int32_t* par = parent + static_cast<size_t>(b) * n;
    int p[4];
    const bool vec = i4 + 3 < n && ((static_cast<size_t>(b) * n) & 3) == 0;
    if (vec) {
        const int4 v = *reinterpret_cast<const int4*>(par + i4);
        p[0] = v.x, p[1] = v.y, p[2] = v.z, p[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = i4 + k < n ? par[i4 + k] : -1;
    }
    uint8_t root[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (p[k] < 0) continue;
        const int a = static_cast<int>(i4) + k;
        int r = p[k];
        // all unions are finished (previous kernel), so roots are fixed points; concurrent
        // compressions by other threads only replace a parent by one of its ancestors
        while (true) {
            const int up = par[r];
            if (up == r) break;
            r = up;
        }
        if (r != p[k]) par[a] = r;
        root[k] = r == a;
    }
    if (MARK_ROOTS) {
        uint8_t* dst = is_root + static_cast<size_t>(b) * n + i4;
        if (vec) {
            *reinterpret_cast<uchar4*>(dst) = make_uchar4(root[0], root[1], root[2], root[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i4 + k < n) dst[k] = root[k];
        }
    }
}

__device__ __forceinline__ void labels_cell(int x, int y, int b, const float* __restrict__ vx, const float* __restrict__ vy,
                                                const uint8_t* __restrict__ state, const int32_t* __restrict__ parent,
                                                const int32_t* __restrict__ rank, const int32_t* __restrict__ root_rank,
                                                int H, int W, int r, EpsTest eps2, int cap,
                                                int32_t* __restrict__ labels, int32_t* __restrict__ indices) {

    const size_t base = static_cast<size_t>(b) * H * W;
    const int o = y * W + x;
    const uint8_t st = state[base + o];
    if (st == 0) return;
    const int slot = rank[base + o];
    if (slot >= cap) return;
    int root = -1;
    if (st == 2) {
        root = parent[base + o];
    } else {
        const float vx0 = vx[base + o], vy0 = vy[base + o];
        for (int dr = -r; dr <= r; ++dr) {
            int yy = y + dr;
            if (yy < 0 || yy >= H) continue;
            for (int dc = -r; dc <= r; ++dc) {
                int xx = x + dc;
                if (xx < 0 || xx >= W) continue;
                int q = yy * W + xx;
                if (state[base + q] == 2 && within_eps(dr, dc, vx0, vy0, vx[base + q], vy[base + q], eps2)) {
                    int rt = parent[base + q];
                    if (root < 0 || rt < root) root = rt;
                }
            }
        }
    }
    const size_t out = static_cast<size_t>(b) * cap + slot;
    labels[out] = root >= 0 ? root_rank[base + root] : -1;
    indices[2 * out] = y;
    indices[2 * out + 1] = x;
}

__global__ void __launch_bounds__(256) k_labels(const float* __restrict__ vx, const float* __restrict__ vy,
                                                const uint8_t* __restrict__ state, const int32_t* __restrict__ parent,
                                                const int32_t* __restrict__ rank, const int32_t* __restrict__ root_rank,
                                                int H, int W, int r, EpsTest eps2, int cap,
                                                int32_t* __restrict__ labels, int32_t* __restrict__ indices) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x4 >= W) return;
    if (vec4_ok(W, state) && *reinterpret_cast<const unsigned*>(state + (static_cast<size_t>(b) * H + y) * W + x4) == 0u)
        return;  // no valid cell among the four
    for (int j = 0; j < 4 && x4 + j < W; ++j)
        labels_cell(x4 + j, y, b, vx, vy, state, parent, rank, root_rank, H, W, r, eps2, cap, labels, indices);
}

static int dbg_tag(int sub) {
    // DATMO_DBSCAN_SUBTAGS=1 spreads the DBSCAN kernels over the other profiler tags (tools/dbscan_prof.py)
    static const bool on = getenv("DATMO_DBSCAN_SUBTAGS") != nullptr;
    return on ? sub : DATMO_TAG_DBSCAN;
}

}  // namespace

// shared with bev.cu (ROI compaction): exclusive rank of every flagged item of [batch][n]
int datmo_flag_scan(datmo_ctx* h, const uint8_t* flags, int64_t n, int batch, int32_t* block_sums, int32_t* totals,
                    int32_t* rank, int tag, int sparse) {
    int nblk = static_cast<int>(ceil_div64(n, SCAN_ITEMS));
    dim3 g(nblk, batch);
    const int vec = (n & 15) == 0 && (reinterpret_cast<uintptr_t>(flags) & 15) == 0;
    {
        LaunchScope ls(h, tag);
        k_flag_block_sums<<<g, SCAN_THREADS, 0, h->stream>>>(flags, n, nblk, block_sums, vec);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag);
        k_scan_block_sums<<<batch, 256, 0, h->stream>>>(block_sums, nblk, totals);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag);
        k_flag_ranks<<<g, SCAN_THREADS, 0, h->stream>>>(flags, n, nblk, block_sums, rank, sparse, vec);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

namespace {

// ---- cluster summaries ------------------------------------------------------------------------
// Integer moments (n, sum r, sum c, sum rr, sum rc, sum cc) are accumulated exactly in uint64;
// sum vx / sum vy in fp64.  Compact cells are in row-major order, so the lanes of a warp mostly
// share a label: they are combined with match_any + shuffles and one lane per distinct label
// issues the atomics.  acc: [batch][max_clusters][8] 64-bit words.
__global__ void __launch_bounds__(256) k_cluster_accum(const float* __restrict__ vx, const float* __restrict__ vy,
                                                       int H, int W, int cap, const int32_t* __restrict__ n_valid,
                                                       const int32_t* __restrict__ labels,
                                                       const int32_t* __restrict__ indices, int max_clusters,
                                                       unsigned long long* __restrict__ acc,
                                                       unsigned long long* __restrict__ extra) {
    const int b = blockIdx.y;
    const int n = min(n_valid[b], cap);
    const int lane = threadIdx.x & 31;
    const int2* idx2 = reinterpret_cast<const int2*>(indices);
    struct Cell {
        int lab;
        unsigned r, c;
        double fvx, fvy;
    };
    // label and (row, col) of compact cell i; the (row, col) load does not wait for the label
    auto head = [&](int i, int& lab, int2& rc) {
        lab = -1, rc = make_int2(0, 0);
        if (i < n) {
            const size_t o = static_cast<size_t>(b) * cap + i;
            lab = labels[o];
            rc = idx2[o];
            if (lab >= max_clusters) lab = -1;
        }
    };
    auto body = [&](int lab, int2 rc) {
        Cell q{lab, 0u, 0u, 0.0, 0.0};
        if (lab >= 0) {
            q.r = rc.x, q.c = rc.y;
            const size_t p = (static_cast<size_t>(b) * H + q.r) * W + q.c;
            q.fvx = vx[p], q.fvy = vy[p];
        }
        return q;
    };
    // Compact cells are row-major, so a warp holds a few RUNS of equal labels.  Segmented inclusive
    // scan over the runs (5 shuffle steps for any mix of labels); the last lane of a run then holds
    // its totals and issues the atomics.  A label split over several runs simply adds several times.
    auto reduce = [&](const Cell& q) {
        const int lab = q.lab;
        const int prev = __shfl_up_sync(0xffffffffu, lab, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != lab);
        const int seg0 = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));  // first lane of my run
        unsigned sr = q.r, sc = q.c;  // <= 32 * 65535: fits
        unsigned long long srr = static_cast<unsigned long long>(q.r) * q.r,
                           src = static_cast<unsigned long long>(q.r) * q.c,
                           scc = static_cast<unsigned long long>(q.c) * q.c;
        double svx = q.fvx, svy = q.fvy;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned tr = __shfl_up_sync(0xffffffffu, sr, d), tc = __shfl_up_sync(0xffffffffu, sc, d);
            const unsigned long long trr = __shfl_up_sync(0xffffffffu, srr, d), trc = __shfl_up_sync(0xffffffffu, src, d),
                                     tcc = __shfl_up_sync(0xffffffffu, scc, d);
            const double tvx = __shfl_up_sync(0xffffffffu, svx, d), tvy = __shfl_up_sync(0xffffffffu, svy, d);
            if (lane - d >= seg0) sr += tr, sc += tc, srr += trr, src += trc, scc += tcc, svx += tvx, svy += tvy;
        }
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        if (tail && lab >= 0) {
            unsigned long long* a = acc + (static_cast<size_t>(b) * max_clusters + lab) * 8;
            atomicAdd(a + 0, static_cast<unsigned long long>(lane - seg0 + 1));
            atomicAdd(a + 1, static_cast<unsigned long long>(sr));
            atomicAdd(a + 2, static_cast<unsigned long long>(sc));
            unsigned long long* x = extra + (static_cast<size_t>(b) * max_clusters + lab) * 4;
            datmo_fixed_add(a + 3, x + 0, x + 2, svx);
            datmo_fixed_add(a + 4, x + 1, x + 3, svy);
            atomicAdd(a + 5, srr);
            atomicAdd(a + 6, src);
            atomicAdd(a + 7, scc);
        }
    };
    // grid-stride over the frame's cells, ROWS warp-rows of 32 cells in flight per thread: the grid is sized for a
    // typical frame, not for `cap`, and the chain label -> (row, col) -> velocity is three dependent memory
    // round trips — the later rows' loads are issued before the first row is reduced
    constexpr int ROWS = 2;   // 4 rows: 58 registers, slower (0.112 vs 0.101 ms per 32 frames)
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x; base < n; base += ROWS * stride) {
        int lab[ROWS];
        int2 rc[ROWS];
#pragma unroll
        for (int k = 0; k < ROWS; ++k) head(base + k * stride + threadIdx.x, lab[k], rc[k]);
        Cell q[ROWS];
#pragma unroll
        for (int k = 0; k < ROWS; ++k) q[k] = body(lab[k], rc[k]);
#pragma unroll
        for (int k = 0; k < ROWS; ++k)
            if (base + k * stride < n) reduce(q[k]);
    }
}

// in place: 64-bit accumulators -> the 8 doubles of the summary
__global__ void __launch_bounds__(256) k_cluster_finalize(int total, unsigned long long* __restrict__ acc,
                                                          const unsigned long long* __restrict__ extra) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    unsigned long long* a = acc + static_cast<size_t>(i) * 8;
    const unsigned long long* x = extra + static_cast<size_t>(i) * 4;
    double* d = reinterpret_cast<double*>(a);
    const unsigned long long cnt = a[0];
    if (cnt == 0) return;  // all-zero bits are also 0.0
    // sums of vx, vy: 128-bit fixed point + the fp64 accumulator for non-finite input
    const double svx = datmo_fixed_value(a[3], x[0], __longlong_as_double(x[2]));
    const double svy = datmo_fixed_value(a[4], x[1], __longlong_as_double(x[3]));
    const double n = static_cast<double>(cnt);
    const double sr = static_cast<double>(a[1]), sc = static_cast<double>(a[2]);
    const double srr = static_cast<double>(a[5]), src = static_cast<double>(a[6]), scc = static_cast<double>(a[7]);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double mr = sr / n, mc = sc / n;
    // np.cov (ddof 1) from exact integer moments
    const double crr = cnt > 1 ? (srr - sr * sr / n) / (n - 1) : nan;
    const double crc = cnt > 1 ? (src - sr * sc / n) / (n - 1) : nan;
    const double ccc = cnt > 1 ? (scc - sc * sc / n) / (n - 1) : nan;
    d[0] = n;
    d[1] = mr;
    d[2] = mc;
    d[3] = svx / n;
    d[4] = svy / n;
    d[5] = crr;
    d[6] = crc;
    d[7] = ccc;
}

}  // namespace

extern "C" int datmo_dbscan_grid_dev(datmo_handle_t h, const float* vx_f, const float* vy_f, const uint8_t* valid,
                                     int H, int W, int batch, double eps, int min_samples, int cap, int32_t* n_valid,
                                     int32_t* labels, int32_t* indices, int32_t* n_clusters) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, vx_f && vy_f && valid && n_valid && labels && indices, "null pointer");
    DATMO_REQUIRE(h, H >= 1 && W >= 1 && batch >= 1 && cap >= 1, "bad sizes");
    DATMO_REQUIRE(h, static_cast<int64_t>(H) * W < (int64_t(1) << 31), "grid too large for int32 cell indices");
    DATMO_REQUIRE(h, H <= 65535 && batch <= 65535, "H and batch must fit a CUDA grid dimension");
    DATMO_REQUIRE(h, eps >= 0 && eps < 1024 && min_samples >= 1, "eps / min_samples out of range");
    const int64_t n = static_cast<int64_t>(H) * W;
    if (datmo_dbscan_runs_supported(eps)) {
        // the reference's eps (5) and anything else with 1 <= floor(eps) <= 15: row runs on bit planes
        DATMO_TRY(datmo_ws_reserve(h, datmo_dbscan_runs_workspace(H, W, batch)));
        return datmo_dbscan_runs(h, h->ws, vx_f, vy_f, valid, H, W, batch, eps, min_samples, cap, n_valid, labels,
                                 indices, n_clusters, dbg_tag);
    }
    const int nblk = static_cast<int>(ceil_div64(n, SCAN_ITEMS));
    const int r = static_cast<int>(floor(eps));
    EpsTest eps2;
    eps2.e2 = eps * eps;
    eps2.lo = static_cast<float>(eps2.e2 * (1.0 - 2e-6));
    eps2.hi = static_cast<float>(eps2.e2 * (1.0 + 2e-6));
    size_t total;
    uint8_t *state, *is_root;
    int32_t *parent, *rank, *root_rank, *bsum, *ncl;
    for (int pass = 0; pass < 2; ++pass) {
        Bump bump(pass ? h->ws : nullptr);
        state = bump.take<uint8_t>(batch * n);
        is_root = bump.take<uint8_t>(batch * n);
        parent = bump.take<int32_t>(batch * n);
        rank = bump.take<int32_t>(batch * n);
        root_rank = bump.take<int32_t>(batch * n);
        bsum = bump.take<int32_t>(static_cast<size_t>(batch) * nblk);
        ncl = bump.take<int32_t>(batch);
        total = bump.off;
        if (!pass) DATMO_TRY(datmo_ws_reserve(h, total));
    }
    DATMO_TRY(datmo_flag_scan(h, valid, n, batch, bsum, n_valid, rank, dbg_tag(0), 1));
    dim3 g(ceil_div(W, 256), H, batch);
    dim3 g4(ceil_div(W, 1024), H, batch);  // four cells per thread
    dim3 gf(static_cast<unsigned>(ceil_div64(n, 1024)), batch);
    {
        {
            LaunchScope ls(h, dbg_tag(1));
            k_core<<<g4, 256, 0, h->stream>>>(vx_f, vy_f, valid, H, W, r, eps2, min_samples, state, parent);
        }
        DATMO_POST_LAUNCH(h);
        if (r >= 1) {
            // (going straight to k_union_near without this atomics-free pass was measured: the union
            // pass grows from 0.28 to 1.24 ms per 32 pairs)
            {
                LaunchScope ls(h, dbg_tag(2));
                k_link_near<<<g4, 256, 0, h->stream>>>(vx_f, vy_f, state, H, W, r, eps2, parent);
            }
            DATMO_POST_LAUNCH(h);
            {
                LaunchScope ls(h, dbg_tag(3));
                k_flatten<false><<<gf, 256, 0, h->stream>>>(n, parent, is_root);
            }
            DATMO_POST_LAUNCH(h);
            {
                LaunchScope ls(h, dbg_tag(5));
                k_union_near<<<g4, 256, 0, h->stream>>>(vx_f, vy_f, state, H, W, r, eps2, parent);
            }
            DATMO_POST_LAUNCH(h);
            if (r > 1) {
                {
                    LaunchScope ls(h, dbg_tag(3));
                    k_flatten<false><<<gf, 256, 0, h->stream>>>(n, parent, is_root);
                }
                DATMO_POST_LAUNCH(h);
                {
                    LaunchScope ls(h, dbg_tag(4));
                    k_union_far<<<g, 256, 0, h->stream>>>(vx_f, vy_f, state, H, W, r, eps2, parent);
                }
                DATMO_POST_LAUNCH(h);
            }
        }
    }
    {
        LaunchScope ls(h, dbg_tag(3));
        k_flatten<true><<<gf, 256, 0, h->stream>>>(n, parent, is_root);
    }
    DATMO_POST_LAUNCH(h);
    int32_t* ncl_out = n_clusters ? n_clusters : ncl;
    DATMO_TRY(datmo_flag_scan(h, is_root, n, batch, bsum, ncl_out, root_rank, dbg_tag(0), 1));
    {
        LaunchScope ls(h, dbg_tag(6));
        k_labels<<<g4, 256, 0, h->stream>>>(vx_f, vy_f, state, parent, rank, root_rank, H, W, r, eps2, cap, labels,
                                           indices);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

extern "C" int datmo_cluster_summary_dev(datmo_handle_t h, const float* vx_f, const float* vy_f, int H, int W,
                                         int batch, int cap, const int32_t* n_valid, const int32_t* labels,
                                         const int32_t* indices, int max_clusters, double* summary) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, vx_f && vy_f && n_valid && labels && indices && summary, "null pointer");
    DATMO_REQUIRE(h, H >= 1 && W >= 1 && batch >= 1 && cap >= 1 && max_clusters >= 1, "bad sizes");
    DATMO_REQUIRE(h, batch <= 65535, "batch must fit a CUDA grid dimension");
    const size_t total = static_cast<size_t>(batch) * max_clusters;
    // low words of the vx / vy sums and their overflow accumulators: 4 words per cluster in the workspace
    DATMO_TRY(datmo_ws_reserve(h, total * 4 * sizeof(unsigned long long)));
    unsigned long long* extra = reinterpret_cast<unsigned long long*>(h->ws);
    DATMO_CHECK_CUDA(h, cudaMemsetAsync(summary, 0, total * 8 * sizeof(double), h->stream));
    DATMO_CHECK_CUDA(h, cudaMemsetAsync(extra, 0, total * 4 * sizeof(unsigned long long), h->stream));
    {
        LaunchScope ls(h, DATMO_TAG_CLUSTER);
        dim3 g(std::max(1, std::min(ceil_div(cap, 256), ceil_div(16 * h->sm_count, batch))), batch);
        k_cluster_accum<<<g, 256, 0, h->stream>>>(vx_f, vy_f, H, W, cap, n_valid, labels, indices, max_clusters,
                                                  reinterpret_cast<unsigned long long*>(summary), extra);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, DATMO_TAG_CLUSTER);
        k_cluster_finalize<<<ceil_div(static_cast<int>(total), 256), 256, 0, h->stream>>>(
            static_cast<int>(total), reinterpret_cast<unsigned long long*>(summary), extra);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

// ---- packed (row, col) for the device-to-host copy -----------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_pack_indices(const int2* __restrict__ indices,
                                                      const int32_t* __restrict__ n_valid, int cap,
                                                      uint32_t* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= min(n_valid[b], cap)) return;
    const int2 rc = indices[static_cast<size_t>(b) * cap + i];
    packed[static_cast<size_t>(b) * cap + i] = (static_cast<uint32_t>(rc.x) << 16) | static_cast<uint32_t>(rc.y);
}
}  // namespace

extern "C" int datmo_pack_indices_dev(datmo_handle_t h, const int32_t* indices, const int32_t* n_valid, int cap,
                                      int batch, uint32_t* packed) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, indices && n_valid && packed && cap >= 1 && batch >= 1, "bad arguments");
    DATMO_REQUIRE(h, batch <= 65535, "batch must fit a CUDA grid dimension");
    {
        LaunchScope ls(h, DATMO_TAG_CLUSTER);
        dim3 g(ceil_div(cap, 256), batch);
        k_pack_indices<<<g, 256, 0, h->stream>>>(reinterpret_cast<const int2*>(indices), n_valid, cap, packed);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}
