// DBSCAN over the valid BEV cells as a union-find of ROW RUNS on bit-packed grids.  sm_100a.
//
// Replaces dbscan_clustering (Optical_flow/main.py:231-259: sklearn.cluster.DBSCAN over the
// (row, col, vx, vy) features of the valid cells) with sklearn-identical labels, numbering
// included, for 1 <= floor(eps) <= 15 (the reference's eps is 5); dbscan.cu keeps the
// cell-level passes for any other eps.  The rule is restated cell by cell in
// oracle/dbscan_runs_np.py and pinned against live sklearn there.
//
// The grid is held as bit planes, one 32-bit word per 32 consecutive cells of a row
// ([batch][H][Ww] words): valid, core, run-head and root bits.  A warp owns 8 consecutive
// words of one row and skips the empty ones (85 % of them on BEV flow fields), a lane is a
// cell, a neighbour window is two funnel shifts over three words.
//   pack      valid bytes -> valid bits (+ the per-warp counts the rank scan needs)
//   core      a valid cell is core when >= min_samples valid cells (itself included) satisfy
//             d2 = drow^2 + dcol^2 + dvx^2 + dvy^2 <= eps^2 (fp64, that order); rows nearest
//             first, early exit: inside a moving region the cell's own row settles it
//   link      run = maximal chain of horizontally adjacent core cells that are pairwise within
//             eps; its head (first cell = minimum index) is the union-find node
//   union     every pair of runs (A in row y, B in row y - dr) within reach of one another is
//             given to ONE cell of A — the first that sees B in its window: A's head for the
//             runs already in the head's window, else the cell at b0 - reach(dr).  That cell
//             compares the two roots and only if they differ walks the cell pairs of (A, B)
//             until one is within eps.  Rows dr = 0, 1 run first; after a flatten the rows
//             dr = 2 .. floor(eps) find almost every pair already joined (6 000 runs,
//             32 000 run pairs, 6 500 cell tests per 150 000-cell frame, against 9 million
//             candidate cell pairs).
//   flatten   heads point at their root (= minimum core index of the cluster); root bits
//   ranks     label of a root = its rank among the roots = sklearn's cluster number
//   labels    core cell: label of its run's root; border cell: smallest root among the core
//             cells within eps (the cluster whose DFS reaches it first), else -1
#include <cstdlib>

#include "common.cuh"
#include "dbscan_common.cuh"

namespace {

constexpr int RUN_MAX_R = 15;     // a (2 r + 1)-cell window must fit one 32-bit word
constexpr int SEG_WORDS = 8;      // words (256 cells) per warp
constexpr int RUN_WARPS = 4;      // warps per CTA

struct RunGeom {
    int H, W, Ww, nseg;   // words per row, 8-word segments per row
    int r, min_samples;
    int rp[RUN_MAX_R + 1];  // reach: largest |dc| with dr^2 + dc^2 <= eps^2
};

// bits of the columns x - rp .. x + rp of a row (bit j <-> column x - rp + j), x = 32 w + lane;
// columns outside the image read as zero
__device__ __forceinline__ uint32_t win_bits(const uint32_t* __restrict__ row, int Ww, int w, int lane, int rp) {
    const uint32_t wm = w > 0 ? row[w - 1] : 0u, w0 = row[w], wp = w + 1 < Ww ? row[w + 1] : 0u;
    const int s = 32 + lane - rp;  // first bit inside the 96-bit string wm | w0 << 32 | wp << 64
    const uint32_t v = s < 32 ? __funnelshift_r(wm, w0, s) : __funnelshift_r(w0, wp, s - 32);
    return v & ((2u << (2 * rp)) - 1u);
}

// first cell of the run that holds core cell x
__device__ __forceinline__ int head_of(const uint32_t* __restrict__ hrow, int x) {
    int w = x >> 5;
    uint32_t m = hrow[w] & (0xffffffffu >> (31 - (x & 31)));
    while (m == 0u && w > 0) m = hrow[--w];
    return m ? 32 * w + 31 - __clz(m) : 0;
}

// last cell of the run that holds core cell x
__device__ __forceinline__ int run_end(const uint32_t* __restrict__ crow, const uint32_t* __restrict__ hrow, int Ww,
                                       int x) {
    int w = x >> 5;
    const int bit = x & 31;
    uint32_t stop = (~crow[w] | hrow[w]) & (bit == 31 ? 0u : 0xffffffffu << (bit + 1));
    while (stop == 0u) {
        if (++w >= Ww) return 32 * Ww - 1;
        stop = ~crow[w] | hrow[w];
    }
    return 32 * w + __ffs(stop) - 2;
}

// The warp's 8 words; calls f(word index, word) for the non-zero ones (warp-uniform loop).
template <typename F>
__device__ __forceinline__ void for_each_word(uint32_t mine, int wbase, F&& f) {
    unsigned nz = __ballot_sync(0xffffffffu, mine != 0u) & ((1u << SEG_WORDS) - 1u);
    while (nz) {
        const int k = __ffs(nz) - 1;
        nz &= nz - 1;
        f(k, wbase + k, __shfl_sync(0xffffffffu, mine, k));
    }
}

struct WarpPos {
    int lane, seg, wbase, y, b;
    bool live;
};
__device__ __forceinline__ WarpPos warp_pos(const RunGeom& g) {
    WarpPos p;
    p.lane = threadIdx.x & 31;
    p.seg = blockIdx.x * RUN_WARPS + (threadIdx.x >> 5);
    p.wbase = p.seg * SEG_WORDS;
    p.y = blockIdx.y;
    p.b = blockIdx.z;
    p.live = p.seg < g.nseg;
    return p;
}
__device__ __forceinline__ uint32_t seg_word(const uint32_t* __restrict__ row, const RunGeom& g, const WarpPos& p) {
    return (p.lane < SEG_WORDS && p.wbase + p.lane < g.Ww) ? row[p.wbase + p.lane] : 0u;
}

// ---- pack: valid bytes -> bits, per-segment counts ---------------------------------------------
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_pack(const uint8_t* __restrict__ valid, RunGeom g,
                                                             uint32_t* __restrict__ vbits,
                                                             int32_t* __restrict__ seg_count) {
    const WarpPos p = warp_pos(g);
    if (!p.live) return;
    const uint8_t* vrow = valid + (static_cast<size_t>(p.b) * g.H + p.y) * g.W;
    uint32_t mine = 0u;
#pragma unroll
    for (int k = 0; k < SEG_WORDS; ++k) {
        const int x = 32 * (p.wbase + k) + p.lane;
        const uint32_t word = __ballot_sync(0xffffffffu, x < g.W && vrow[x] != 0);
        if (p.lane == k) mine = word;
    }
    if (p.lane < SEG_WORDS && p.wbase + p.lane < g.Ww)
        vbits[(static_cast<size_t>(p.b) * g.H + p.y) * g.Ww + p.wbase + p.lane] = mine;
    const int cnt = __reduce_add_sync(0xffffffffu, __popc(mine));
    if (p.lane == 0) seg_count[(static_cast<size_t>(p.b) * g.H + p.y) * g.nseg + p.seg] = cnt;
}

// ---- scan of the per-segment counts (one CTA per frame), optionally numbering the roots --------
// counts -> exclusive offsets in place; totals[b] = sum.  With rbits: every root cell's label
// (its rank among the frame's roots in row-major order) is written to rlabel[cell].
__global__ void __launch_bounds__(256) k_run_scan(int32_t* __restrict__ seg_count, int nblk, int32_t* __restrict__ totals,
                                                  const uint32_t* __restrict__ rbits, RunGeom g,
                                                  int32_t* __restrict__ rlabel) {
    const int b = blockIdx.x;
    int32_t* s = seg_count + static_cast<size_t>(b) * nblk;
    __shared__ int s_part[256];
    const int per = (nblk + 255) / 256;
    const int lo = min(threadIdx.x * per, nblk), hi = min(lo + per, nblk);
    int t = 0;
    for (int i = lo; i < hi; ++i) t += s[i];
    // CTA-wide exclusive scan of t
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_part[wid] = inc;
    __syncthreads();
    int woff = 0, all = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < wid) woff += s_part[i];
        all += s_part[i];
    }
    if (threadIdx.x == 0) totals[b] = all;
    int run = woff + inc - t;
    for (int i = lo; i < hi; ++i) {
        const int v = s[i];
        s[i] = run;
        if (rbits != nullptr && v > 0) {
            // number the roots of this segment
            const int y = i / g.nseg, seg = i - y * g.nseg;
            const uint32_t* row = rbits + (static_cast<size_t>(b) * g.H + y) * g.Ww;
            int k = run;
            for (int w = seg * SEG_WORDS; w < min((seg + 1) * SEG_WORDS, g.Ww); ++w) {
                uint32_t bits = row[w];
                while (bits) {
                    const int bit = __ffs(bits) - 1;
                    bits &= bits - 1;
                    rlabel[static_cast<size_t>(b) * g.H * g.W + static_cast<size_t>(y) * g.W + 32 * w + bit] = k++;
                }
            }
        }
        run += v;
    }
}

// ---- core cells ------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_core(const float* __restrict__ vx, const float* __restrict__ vy,
                                                             const uint32_t* __restrict__ vbits, RunGeom g,
                                                             EpsTest eps2, uint32_t* __restrict__ cbits) {
    const WarpPos p = warp_pos(g);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const uint32_t* vimg = vbits + img * g.Ww;
    const float* vxi = vx + img * g.W;
    const float* vyi = vy + img * g.W;
    const uint32_t mine = seg_word(vimg + static_cast<size_t>(p.y) * g.Ww, g, p);
    uint32_t cmine = 0u;
    for_each_word(mine, p.wbase, [&](int k, int w, uint32_t word) {
        bool core = false;
        if ((word >> p.lane) & 1u) {
            const int x = 32 * w + p.lane;
            const float vx0 = vxi[static_cast<size_t>(p.y) * g.W + x], vy0 = vyi[static_cast<size_t>(p.y) * g.W + x];
            int cnt = 0;
            // rows in the order 0, -1, +1, -2, +2, ..: inside a moving region the own row suffices
            for (int i = 0; i <= 2 * g.r && cnt < g.min_samples; ++i) {
                const int dr = (i & 1) ? -((i + 1) >> 1) : (i >> 1);
                const int yy = p.y + dr;
                if (yy < 0 || yy >= g.H) continue;
                const int rp = g.rp[dr < 0 ? -dr : dr];
                uint32_t win = win_bits(vimg + static_cast<size_t>(yy) * g.Ww, g.Ww, w, p.lane, rp);
                const float* nvx = vxi + static_cast<size_t>(yy) * g.W + (x - rp);
                const float* nvy = vyi + static_cast<size_t>(yy) * g.W + (x - rp);
                while (win) {
                    const int j = __ffs(win) - 1;
                    win &= win - 1;
                    if (within_eps(dr, j - rp, vx0, vy0, nvx[j], nvy[j], eps2) && ++cnt >= g.min_samples) break;
                }
            }
            core = cnt >= g.min_samples;
        }
        const uint32_t cword = __ballot_sync(0xffffffffu, core);
        if (p.lane == k) cmine = cword;
    });
    if (p.lane < SEG_WORDS && p.wbase + p.lane < g.Ww) cbits[(img + p.y) * g.Ww + p.wbase + p.lane] = cmine;
}

// ---- runs: head bits, parent[head] = head ---------------------------------------------------------
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_link(const float* __restrict__ vx, const float* __restrict__ vy,
                                                             const uint32_t* __restrict__ cbits, RunGeom g,
                                                             EpsTest eps2, uint32_t* __restrict__ hbits,
                                                             int32_t* __restrict__ parent) {
    const WarpPos p = warp_pos(g);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const uint32_t* crow = cbits + (img + p.y) * g.Ww;
    const float* vxr = vx + (img + p.y) * g.W;
    const float* vyr = vy + (img + p.y) * g.W;
    int32_t* par = parent + img * g.W;
    const uint32_t mine = seg_word(crow, g, p);
    uint32_t hmine = 0u;
    for_each_word(mine, p.wbase, [&](int k, int w, uint32_t word) {
        const bool core = (word >> p.lane) & 1u;
        const bool left = p.lane > 0 ? (word >> (p.lane - 1)) & 1u : (w > 0 && (crow[w - 1] >> 31));
        const int x = 32 * w + p.lane;
        bool link = false;
        if (core && left) link = within_eps(0, 1, vxr[x], vyr[x], vxr[x - 1], vyr[x - 1], eps2);
        const uint32_t hword = word & ~__ballot_sync(0xffffffffu, link);
        if (p.lane == k) hmine = hword;
        if (core && !link) par[p.y * g.W + x] = p.y * g.W + x;
    });
    if (p.lane < SEG_WORDS && p.wbase + p.lane < g.Ww) hbits[(img + p.y) * g.Ww + p.wbase + p.lane] = hmine;
}

// ---- unions between runs -------------------------------------------------------------------------
// Run of (y, x) against the run of row yy = y - dr that holds column xb; (y, x) is the cell
// responsible for the pair.  a0 / b0: the runs' heads when the caller knows them, else -1.
__device__ __noinline__ void run_pair(const RunGeom& g, const EpsTest& eps2, const float* __restrict__ vxi,
                                      const float* __restrict__ vyi, const uint32_t* __restrict__ cimg,
                                      const uint32_t* __restrict__ himg, int32_t* par, int y, int x, int yy, int xb,
                                      int a0, int b0) {
    const uint32_t* hrow_a = himg + static_cast<size_t>(y) * g.Ww;
    const uint32_t* hrow_b = himg + static_cast<size_t>(yy) * g.Ww;
    if (a0 < 0) a0 = head_of(hrow_a, x);
    if (b0 < 0) b0 = head_of(hrow_b, xb);
    const int ra = uf_find(par, y * g.W + a0), rb = uf_find(par, yy * g.W + b0);
    if (ra == rb) return;
    const int a1 = run_end(cimg + static_cast<size_t>(y) * g.Ww, hrow_a, g.Ww, x);
    const int b1 = run_end(cimg + static_cast<size_t>(yy) * g.Ww, hrow_b, g.Ww, b0);
    const int dr = y - yy, rp = g.rp[dr];
    const float* ax = vxi + static_cast<size_t>(y) * g.W;
    const float* ay = vyi + static_cast<size_t>(y) * g.W;
    const float* bx = vxi + static_cast<size_t>(yy) * g.W;
    const float* by = vyi + static_cast<size_t>(yy) * g.W;
    for (int xa = x; xa <= a1 && xa - rp <= b1; ++xa) {
        const int lo = max(b0, xa - rp), hi = dr > 0 ? min(b1, xa + rp) : min(b1, xa - 1);
        const float vx0 = ax[xa], vy0 = ay[xa];
        for (int q = lo; q <= hi; ++q) {
            if (within_eps(dr, xa - q, vx0, vy0, bx[q], by[q], eps2)) {
                uf_union(par, ra, rb);
                return;
            }
        }
    }
}

__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_union(const float* __restrict__ vx, const float* __restrict__ vy,
                                                              const uint32_t* __restrict__ cbits,
                                                              const uint32_t* __restrict__ hbits, RunGeom g,
                                                              EpsTest eps2, int dr_lo, int dr_hi,
                                                              int32_t* __restrict__ parent) {
    const WarpPos p = warp_pos(g);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const uint32_t* cimg = cbits + img * g.Ww;
    const uint32_t* himg = hbits + img * g.Ww;
    const float* vxi = vx + img * g.W;
    const float* vyi = vy + img * g.W;
    int32_t* par = parent + img * g.W;
    const uint32_t mine = seg_word(cimg + static_cast<size_t>(p.y) * g.Ww, g, p);
    for_each_word(mine, p.wbase, [&](int, int w, uint32_t word) {
        if (!((word >> p.lane) & 1u)) return;
        const int x = 32 * w + p.lane;
        const bool head = (himg[static_cast<size_t>(p.y) * g.Ww + w] >> p.lane) & 1u;
        for (int dr = dr_lo; dr <= dr_hi; ++dr) {
            const int yy = p.y - dr;
            if (yy < 0) break;
            const int rp = g.rp[dr];
            const uint32_t hwin = win_bits(himg + static_cast<size_t>(yy) * g.Ww, g.Ww, w, p.lane, rp);
            if (head) {
                uint32_t cwin = win_bits(cimg + static_cast<size_t>(yy) * g.Ww, g.Ww, w, p.lane, rp);
                if (dr == 0) cwin &= (1u << rp) - 1u;  // own row: the columns x - r .. x - 1
                // a run starts at every head bit and at the window's first core cell
                uint32_t starts = (hwin & cwin) | (cwin & (0u - cwin));
                while (starts) {
                    const int j = __ffs(starts) - 1;
                    starts &= starts - 1;
                    const int xb = x - rp + j;
                    run_pair(g, eps2, vxi, vyi, cimg, himg, par, p.y, x, yy, xb, x, ((hwin >> j) & 1u) ? xb : -1);
                }
            } else if (dr > 0 && ((hwin >> (2 * rp)) & 1u)) {
                // a run of the row above whose head enters the window at its right edge
                run_pair(g, eps2, vxi, vyi, cimg, himg, par, p.y, x, yy, x + rp, -1, x + rp);
            }
        }
    });
}

// ---- flatten: heads point at their root; MARK: root bits + per-segment root counts ----------------
template <bool MARK>
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_flatten(const uint32_t* __restrict__ hbits, RunGeom g,
                                                                int32_t* __restrict__ parent,
                                                                uint32_t* __restrict__ rbits,
                                                                int32_t* __restrict__ seg_count) {
    const WarpPos p = warp_pos(g);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    int32_t* par = parent + img * g.W;
    const uint32_t mine = seg_word(hbits + (img + p.y) * g.Ww, g, p);
    uint32_t rmine = 0u;
    for_each_word(mine, p.wbase, [&](int k, int w, uint32_t word) {
        bool root = false;
        if ((word >> p.lane) & 1u) {
            const int a = p.y * g.W + 32 * w + p.lane;
            // every union is finished (previous kernel): roots are fixed points, and concurrent
            // compressions only replace a parent by one of its ancestors
            const int first = __ldcg(par + a);
            int r = first;
            while (true) {
                const int up = __ldcg(par + r);
                if (up == r) break;
                r = up;
            }
            if (r != first) __stcg(par + a, r);
            root = r == a;
        }
        if (MARK) {
            const uint32_t rword = __ballot_sync(0xffffffffu, root);
            if (p.lane == k) rmine = rword;
        }
    });
    if (MARK) {
        if (p.lane < SEG_WORDS && p.wbase + p.lane < g.Ww) rbits[(img + p.y) * g.Ww + p.wbase + p.lane] = rmine;
        const int cnt = __reduce_add_sync(0xffffffffu, __popc(rmine));
        if (p.lane == 0) seg_count[(img + p.y) * g.nseg + p.seg] = cnt;
    }
}

// ---- labels ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_labels(const float* __restrict__ vx, const float* __restrict__ vy,
                                                               const uint32_t* __restrict__ vbits,
                                                               const uint32_t* __restrict__ cbits,
                                                               const uint32_t* __restrict__ hbits, RunGeom g,
                                                               EpsTest eps2, const int32_t* __restrict__ parent,
                                                               const int32_t* __restrict__ rlabel,
                                                               const int32_t* __restrict__ seg_off, int cap,
                                                               int32_t* __restrict__ labels,
                                                               int32_t* __restrict__ indices) {
    const WarpPos p = warp_pos(g);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const uint32_t* vimg = vbits + img * g.Ww;
    const uint32_t* cimg = cbits + img * g.Ww;
    const uint32_t* himg = hbits + img * g.Ww;
    const float* vxi = vx + img * g.W;
    const float* vyi = vy + img * g.W;
    const int32_t* par = parent + img * g.W;
    const int32_t* rl = rlabel + img * g.W;
    const uint32_t mine = seg_word(vimg + static_cast<size_t>(p.y) * g.Ww, g, p);
    if (__ballot_sync(0xffffffffu, mine != 0u) == 0u) return;
    // row-major rank of the first valid cell of every word of the segment
    int before = __popc(mine);
#pragma unroll
    for (int o = 1; o < SEG_WORDS; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, before, o);
        if (p.lane >= o) before += v;
    }
    before += seg_off[(img + p.y) * g.nseg + p.seg] - __popc(mine);
    for_each_word(mine, p.wbase, [&](int k, int w, uint32_t word) {
        const int base = __shfl_sync(0xffffffffu, before, k);
        if (!((word >> p.lane) & 1u)) return;
        const int slot = base + __popc(word & ((1u << p.lane) - 1u));
        if (slot >= cap) return;
        const int x = 32 * w + p.lane;
        int root = -1;
        if ((cimg[static_cast<size_t>(p.y) * g.Ww + w] >> p.lane) & 1u) {
            root = par[p.y * g.W + head_of(himg + static_cast<size_t>(p.y) * g.Ww, x)];
        } else {
            const float vx0 = vxi[static_cast<size_t>(p.y) * g.W + x], vy0 = vyi[static_cast<size_t>(p.y) * g.W + x];
            for (int dr = -g.r; dr <= g.r; ++dr) {
                const int yy = p.y + dr;
                if (yy < 0 || yy >= g.H) continue;
                const int rp = g.rp[dr < 0 ? -dr : dr];
                uint32_t win = win_bits(cimg + static_cast<size_t>(yy) * g.Ww, g.Ww, w, p.lane, rp);
                while (win) {
                    const int j = __ffs(win) - 1;
                    win &= win - 1;
                    const int xx = x - rp + j;
                    if (within_eps(dr, j - rp, vx0, vy0, vxi[static_cast<size_t>(yy) * g.W + xx],
                                   vyi[static_cast<size_t>(yy) * g.W + xx], eps2)) {
                        const int rt = par[yy * g.W + head_of(himg + static_cast<size_t>(yy) * g.Ww, xx)];
                        if (root < 0 || rt < root) root = rt;
                    }
                }
            }
        }
        const size_t out = static_cast<size_t>(p.b) * cap + slot;
        labels[out] = root >= 0 ? rl[root] : -1;
        reinterpret_cast<int2*>(indices)[out] = make_int2(p.y, x);
    });
}

}  // namespace

bool datmo_dbscan_runs_supported(double eps) {
    const bool off = getenv("DATMO_DBSCAN_CELLS") != nullptr;  // A/B runs against the cell-level passes
    const int r = static_cast<int>(floor(eps));
    return !off && r >= 1 && r <= RUN_MAX_R;
}

size_t datmo_dbscan_runs_workspace(int H, int W, int batch) {
    const size_t Ww = (W + 31) / 32, nseg = (Ww + SEG_WORDS - 1) / SEG_WORDS;
    Bump bump(nullptr);
    for (int i = 0; i < 4; ++i) bump.take<uint32_t>(static_cast<size_t>(batch) * H * Ww);
    for (int i = 0; i < 2; ++i) bump.take<int32_t>(static_cast<size_t>(batch) * H * nseg);
    for (int i = 0; i < 2; ++i) bump.take<int32_t>(static_cast<size_t>(batch) * H * W);
    bump.take<int32_t>(batch);
    return bump.off;
}

// Called by datmo_dbscan_grid_dev (dbscan.cu) with validated arguments; ws holds
// datmo_dbscan_runs_workspace bytes.  tag(i) maps a sub-stage to a profiler tag.
int datmo_dbscan_runs(datmo_ctx* h, char* ws, const float* vx_f, const float* vy_f, const uint8_t* valid, int H, int W,
                      int batch, double eps, int min_samples, int cap, int32_t* n_valid, int32_t* labels,
                      int32_t* indices, int32_t* n_clusters, int (*tag)(int)) {
    RunGeom g;
    g.H = H, g.W = W, g.Ww = (W + 31) / 32, g.nseg = (g.Ww + SEG_WORDS - 1) / SEG_WORDS;
    g.r = static_cast<int>(floor(eps));
    g.min_samples = min_samples;
    EpsTest eps2;
    eps2.e2 = eps * eps;
    eps2.lo = static_cast<float>(eps2.e2 * (1.0 - 2e-6));
    eps2.hi = static_cast<float>(eps2.e2 * (1.0 + 2e-6));
    for (int dr = 0; dr <= RUN_MAX_R; ++dr) {
        int dc = 0;
        while (dr <= g.r && static_cast<double>(dr * dr + (dc + 1) * (dc + 1)) <= eps2.e2) ++dc;
        g.rp[dr] = dc;
    }
    Bump bump(ws);
    const size_t nw = static_cast<size_t>(batch) * H * g.Ww, ns = static_cast<size_t>(batch) * H * g.nseg;
    uint32_t* vbits = bump.take<uint32_t>(nw);
    uint32_t* cbits = bump.take<uint32_t>(nw);
    uint32_t* hbits = bump.take<uint32_t>(nw);
    uint32_t* rbits = bump.take<uint32_t>(nw);
    int32_t* vseg = bump.take<int32_t>(ns);
    int32_t* rseg = bump.take<int32_t>(ns);
    int32_t* parent = bump.take<int32_t>(static_cast<size_t>(batch) * H * W);
    int32_t* rlabel = bump.take<int32_t>(static_cast<size_t>(batch) * H * W);
    int32_t* ncl = bump.take<int32_t>(batch);
    const dim3 grid(ceil_div(g.nseg, RUN_WARPS), H, batch);
    const int nt = 32 * RUN_WARPS, nblk = H * g.nseg;
    cudaStream_t s = h->stream;
    {
        LaunchScope ls(h, tag(0));
        k_run_pack<<<grid, nt, 0, s>>>(valid, g, vbits, vseg);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(0));
        k_run_scan<<<batch, 256, 0, s>>>(vseg, nblk, n_valid, nullptr, g, nullptr);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(1));
        k_run_core<<<grid, nt, 0, s>>>(vx_f, vy_f, vbits, g, eps2, cbits);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(2));
        k_run_link<<<grid, nt, 0, s>>>(vx_f, vy_f, cbits, g, eps2, hbits, parent);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(5));
        k_run_union<<<grid, nt, 0, s>>>(vx_f, vy_f, cbits, hbits, g, eps2, 0, 1, parent);
    }
    DATMO_POST_LAUNCH(h);
    if (g.r > 1) {
        {
            LaunchScope ls(h, tag(3));
            k_run_flatten<false><<<grid, nt, 0, s>>>(hbits, g, parent, nullptr, nullptr);
        }
        DATMO_POST_LAUNCH(h);
        {
            LaunchScope ls(h, tag(4));
            k_run_union<<<grid, nt, 0, s>>>(vx_f, vy_f, cbits, hbits, g, eps2, 2, g.r, parent);
        }
        DATMO_POST_LAUNCH(h);
    }
    {
        LaunchScope ls(h, tag(3));
        k_run_flatten<true><<<grid, nt, 0, s>>>(hbits, g, parent, rbits, rseg);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(0));
        k_run_scan<<<batch, 256, 0, s>>>(rseg, nblk, n_clusters ? n_clusters : ncl, rbits, g, rlabel);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(6));
        k_run_labels<<<grid, nt, 0, s>>>(vx_f, vy_f, vbits, cbits, hbits, g, eps2, parent, rlabel, vseg, cap, labels,
                                         indices);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}
