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
constexpr int SEG_WORDS = 32;     // words (1024 cells) of one row per warp: one coalesced 128-byte load
constexpr int RUN_WARPS = 4;      // warps per CTA

struct RunGeom {
    int H, W, Ww, nseg;   // words per row, 32-word segments per row
    int r, min_samples;
    int rp[RUN_MAX_R + 1];  // reach: largest |dc| with dr^2 + dc^2 <= eps^2
};

// three consecutive words of a bit-plane row around word w (zero outside the row)
struct W3 {
    uint32_t m, c, p;
};
__device__ __forceinline__ W3 load3(const uint32_t* __restrict__ row, int Ww, int w) {
    W3 t;
    t.m = w > 0 ? row[w - 1] : 0u;
    t.c = row[w];
    t.p = w + 1 < Ww ? row[w + 1] : 0u;
    return t;
}
// bits of the columns x - rp .. x + rp (bit j <-> column x - rp + j), x = bit `pos` of the centre word
__device__ __forceinline__ uint32_t win_bits(const W3& t, int pos, int rp) {
    const int s = 32 + pos - rp;  // first bit inside the 96-bit string m | c << 32 | p << 64
    const uint32_t v = s < 32 ? __funnelshift_r(t.m, t.c, s) : __funnelshift_r(t.c, t.p, s - 32);
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

// Warp = one 32-word segment of one row.  lane <-> word for the bit-only work; for the per-cell work
// for_each_word walks the non-zero words (warp-uniform loop) with lane <-> cell.
template <typename F>
__device__ __forceinline__ void for_each_word(uint32_t mine, int wbase, F&& f) {
    unsigned nz = __ballot_sync(0xffffffffu, mine != 0u);
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
// grid (ceil(rows * nseg / RUN_WARPS), batch): consecutive warps take consecutive (row, segment) pairs
__device__ __forceinline__ WarpPos warp_pos(const RunGeom& g, int rows) {
    WarpPos p;
    p.lane = threadIdx.x & 31;
    const int gw = blockIdx.x * RUN_WARPS + (threadIdx.x >> 5);
    p.y = gw / g.nseg;
    p.seg = gw - p.y * g.nseg;
    p.wbase = p.seg * SEG_WORDS;
    p.b = blockIdx.y;
    p.live = p.y < rows;
    return p;
}
__device__ __forceinline__ uint32_t seg_word(const uint32_t* __restrict__ row, const RunGeom& g, const WarpPos& p) {
    return p.wbase + p.lane < g.Ww ? row[p.wbase + p.lane] : 0u;
}

// ---- pack: valid bytes -> bits, per-segment counts ---------------------------------------------
__device__ __forceinline__ uint32_t nonzero_bytes16(const uint4 v) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    uint32_t bits = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned m = __vcmpne4(w[j], 0u);  // 0xff per non-zero byte
        bits |= ((m & 1u) | ((m >> 7) & 2u) | ((m >> 14) & 4u) | ((m >> 21) & 8u)) << (4 * j);
    }
    return bits;
}

__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_pack(const uint8_t* __restrict__ valid, RunGeom g, int vec,
                                                             uint32_t* __restrict__ vbits,
                                                             int32_t* __restrict__ seg_count) {
    const WarpPos p = warp_pos(g, g.H);
    if (!p.live) return;
    const uint8_t* vrow = valid + (static_cast<size_t>(p.b) * g.H + p.y) * g.W;
    const int w = p.wbase + p.lane, x0 = 32 * w;
    uint32_t mine = 0u;
    if (w < g.Ww) {
        if (vec && x0 + 32 <= g.W) {
            // the lane's 32 cells as two 16-byte loads (a warp reads 1 KiB of the row)
            const uint4* q = reinterpret_cast<const uint4*>(vrow + x0);
            mine = nonzero_bytes16(q[0]) | (nonzero_bytes16(q[1]) << 16);
        } else {
            for (int i = 0; i < 32 && x0 + i < g.W; ++i) mine |= (vrow[x0 + i] != 0 ? 1u : 0u) << i;
        }
        vbits[(static_cast<size_t>(p.b) * g.H + p.y) * g.Ww + w] = mine;
    }
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
    __shared__ int s_part[8];
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

// ---- core cells + horizontal links ---------------------------------------------------------------
// cbits: core cells.  pbits: valid cells whose left neighbour is valid and within eps (a run link
// once both turn out to be core).  A warp walks the non-empty words of its row segment with
// lane <-> cell; the velocities of the NEXT word are in flight while the current one is counted, and
// the own row's window comes out of the words the warp already holds.
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_core(const float* __restrict__ vx, const float* __restrict__ vy,
                                                             const uint32_t* __restrict__ vbits, RunGeom g,
                                                             EpsTest eps2, uint32_t* __restrict__ cbits,
                                                             uint32_t* __restrict__ pbits) {
    const WarpPos p = warp_pos(g, g.H);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const uint32_t* vimg = vbits + img * g.Ww;
    const uint32_t* vrow = vimg + static_cast<size_t>(p.y) * g.Ww;
    const float* vxi = vx + img * g.W;
    const float* vyi = vy + img * g.W;
    const float* rx = vxi + static_cast<size_t>(p.y) * g.W;
    const float* ry = vyi + static_cast<size_t>(p.y) * g.W;
    const uint32_t mine = seg_word(vrow, g, p);
    uint32_t cmine = 0u, pmine = 0u;
    unsigned nz = __ballot_sync(0xffffffffu, mine != 0u);
    if (nz) {
        const uint32_t edge_l = p.wbase > 0 ? vrow[p.wbase - 1] : 0u;
        const uint32_t edge_r = p.wbase + SEG_WORDS < g.Ww ? vrow[p.wbase + SEG_WORDS] : 0u;
        int k = __ffs(nz) - 1;
        uint32_t word = __shfl_sync(0xffffffffu, mine, k);
        float vx0 = 0.f, vy0 = 0.f;
        if ((word >> p.lane) & 1u) vx0 = rx[32 * (p.wbase + k) + p.lane], vy0 = ry[32 * (p.wbase + k) + p.lane];
        while (true) {
            nz &= nz - 1;
            const int kn = nz ? __ffs(nz) - 1 : -1;
            uint32_t word_n = 0u;
            float vx_n = 0.f, vy_n = 0.f;
            if (kn >= 0) {
                word_n = __shfl_sync(0xffffffffu, mine, kn);
                if ((word_n >> p.lane) & 1u)
                    vx_n = rx[32 * (p.wbase + kn) + p.lane], vy_n = ry[32 * (p.wbase + kn) + p.lane];
            }
            const int w = p.wbase + k, x = 32 * w + p.lane;
            W3 own;
            own.c = word;
            own.m = __shfl_sync(0xffffffffu, mine, max(k - 1, 0));
            own.p = __shfl_sync(0xffffffffu, mine, min(k + 1, 31));
            if (k == 0) own.m = edge_l;
            if (k == 31) own.p = edge_r;
            const bool valid = (word >> p.lane) & 1u;
            float lvx = __shfl_up_sync(0xffffffffu, vx0, 1), lvy = __shfl_up_sync(0xffffffffu, vy0, 1);
            bool core = false, plink = false;
            if (valid) {
                const bool left = p.lane > 0 ? (word >> (p.lane - 1)) & 1u : (own.m >> 31);
                if (left) {
                    if (p.lane == 0) lvx = rx[x - 1], lvy = ry[x - 1];
                    plink = within_eps(0, 1, vx0, vy0, lvx, lvy, eps2);
                }
                int cnt = 0;
                // rows in the order 0, -1, +1, -2, +2, ..: inside a moving region the own row suffices
                for (int i = 0; i <= 2 * g.r && cnt < g.min_samples; ++i) {
                    const int dr = (i & 1) ? -((i + 1) >> 1) : (i >> 1);
                    const int yy = p.y + dr;
                    if (yy < 0 || yy >= g.H) continue;
                    const int rp = g.rp[dr < 0 ? -dr : dr];
                    uint32_t win = win_bits(i == 0 ? own : load3(vimg + static_cast<size_t>(yy) * g.Ww, g.Ww, w), p.lane, rp);
                    const float* nvx = vxi + static_cast<size_t>(yy) * g.W + (x - rp);
                    const float* nvy = vyi + static_cast<size_t>(yy) * g.W + (x - rp);
                    while (win && cnt < g.min_samples) {
                        // up to four candidates' velocities in flight together
                        int j[4];
                        float cx[4], cy[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            j[u] = win ? __ffs(win) - 1 : -1;
                            win &= win - 1;   // 0 & anything stays 0
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (j[u] >= 0) cx[u] = nvx[j[u]], cy[u] = nvy[j[u]];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (j[u] >= 0 && within_eps(dr, j[u] - rp, vx0, vy0, cx[u], cy[u], eps2)) ++cnt;
                    }
                }
                core = cnt >= g.min_samples;
            }
            const uint32_t cword = __ballot_sync(0xffffffffu, core), pword = __ballot_sync(0xffffffffu, plink);
            if (p.lane == k) cmine = cword, pmine = pword;
            if (kn < 0) break;
            k = kn, word = word_n, vx0 = vx_n, vy0 = vy_n;
        }
    }
    if (p.wbase + p.lane < g.Ww) {
        cbits[(img + p.y) * g.Ww + p.wbase + p.lane] = cmine;
        pbits[(img + p.y) * g.Ww + p.wbase + p.lane] = pmine;
    }
}

// ---- runs: head bits, parent[head] = head (bit operations only; lane <-> word) ---------------------
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_heads(const uint32_t* __restrict__ cbits,
                                                              const uint32_t* __restrict__ pbits, RunGeom g,
                                                              uint32_t* __restrict__ hbits,
                                                              int32_t* __restrict__ parent) {
    const WarpPos p = warp_pos(g, g.H);
    if (!p.live) return;
    const int w = p.wbase + p.lane;
    if (w >= g.Ww) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const uint32_t* crow = cbits + (img + p.y) * g.Ww;
    const uint32_t c = crow[w];
    const uint32_t left = (c << 1) | (w > 0 ? crow[w - 1] >> 31 : 0u);
    uint32_t h = c & ~(pbits[(img + p.y) * g.Ww + w] & left);
    hbits[(img + p.y) * g.Ww + w] = h;
    int32_t* par = parent + img * g.W;
    while (h) {
        const int a = p.y * g.W + 32 * w + __ffs(h) - 1;
        h &= h - 1;
        par[a] = a;
    }
}

// ---- unions between runs -------------------------------------------------------------------------
// link root ra and root rb (ra != rb when called; both were roots a moment ago)
__device__ __forceinline__ void uf_link(int32_t* parent, int a, int b) {
    while (true) {
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicCAS(parent + a, a, b);
        if (old == a) return;
        a = uf_find(parent, old);
        b = uf_find(parent, b);
    }
}

// Run of (y, x) against the run of row yy = y - dr whose first cell inside the window is column xb;
// (y, x) is the cell responsible for the pair.  a0 / b0: the runs' heads when the caller knows them,
// else -1.  The two parents and the velocities of the first cell pair (x, xb) are fetched together:
// one memory round trip decides "already joined" (QUICK: the forest was flattened before this kernel,
// so equal parents prove it) and, when the first pair is within eps, the union follows at once.  Only
// when that pair fails are the run ends looked up and the remaining cell pairs of (A, B) walked.
template <bool QUICK>
__device__ __forceinline__ void run_pair(const RunGeom& g, const EpsTest& eps2, const float* __restrict__ vxi,
                                         const float* __restrict__ vyi, const uint32_t* __restrict__ cimg,
                                         const uint32_t* __restrict__ himg, int32_t* par, int y, int x, int yy, int xb,
                                         int a0, int b0) {
    const uint32_t* hrow_a = himg + static_cast<size_t>(y) * g.Ww;
    const uint32_t* hrow_b = himg + static_cast<size_t>(yy) * g.Ww;
    const float* ax = vxi + static_cast<size_t>(y) * g.W;
    const float* ay = vyi + static_cast<size_t>(y) * g.W;
    const float* bx = vxi + static_cast<size_t>(yy) * g.W;
    const float* by = vyi + static_cast<size_t>(yy) * g.W;
    const float vxa = ax[x], vya = ay[x], vxb = bx[xb], vyb = by[xb];
    if (a0 < 0) a0 = head_of(hrow_a, x);
    if (b0 < 0) b0 = head_of(hrow_b, xb);
    int ra = y * g.W + a0, rb = yy * g.W + b0;
    const int pa = __ldcg(par + ra), pb = __ldcg(par + rb);
    if (QUICK && pa == pb) return;
    if (pa != ra) ra = uf_find(par, pa);
    if (pb != rb) rb = uf_find(par, pb);
    if (ra == rb) return;
    const int dr = y - yy, rp = g.rp[dr];
    if (within_eps(dr, x - xb, vxa, vya, vxb, vyb, eps2)) {
        uf_link(par, ra, rb);
        return;
    }
    const int a1 = run_end(cimg + static_cast<size_t>(y) * g.Ww, hrow_a, g.Ww, x);
    const int b1 = run_end(cimg + static_cast<size_t>(yy) * g.Ww, hrow_b, g.Ww, b0);
    for (int xa = x; xa <= a1 && xa - rp <= b1; ++xa) {
        const int lo = max(b0, xa - rp), hi = dr > 0 ? min(b1, xa + rp) : min(b1, xa - 1);
        const float vx0 = ax[xa], vy0 = ay[xa];
        for (int q = lo; q <= hi; ++q) {
            if (within_eps(dr, xa - q, vx0, vy0, bx[q], by[q], eps2)) {
                uf_link(par, ra, rb);
                return;
            }
        }
    }
}

// All run pairs the cells of word w of row y are responsible for, against rows y - DR_LO .. y - dr_hi
// (not above row y_min).  One thread per word: the responsible cells come out of bit operations.  The
// words of every row involved are fetched before any is looked at (independent loads).
template <bool QUICK, int DR_LO, int NROWS>
__device__ __forceinline__ void word_pairs(const RunGeom& g, const EpsTest& eps2, const float* __restrict__ vxi,
                                           const float* __restrict__ vyi, const uint32_t* __restrict__ cimg,
                                           const uint32_t* __restrict__ himg, int32_t* par, int y, int w, int dr_hi,
                                           int y_min) {
    const uint32_t c = cimg[static_cast<size_t>(y) * g.Ww + w];
    if (c == 0u) return;
    const uint32_t hd = himg[static_cast<size_t>(y) * g.Ww + w];
    W3 ht[NROWS], ct[NROWS];
#pragma unroll
    for (int i = 0; i < NROWS; ++i) {
        const int yy = y - (DR_LO + i);
        const bool on = DR_LO + i <= dr_hi && yy >= y_min;
        ht[i] = on ? load3(himg + static_cast<size_t>(yy) * g.Ww, g.Ww, w) : W3{0u, 0u, 0u};
        ct[i] = on && hd ? load3(cimg + static_cast<size_t>(yy) * g.Ww, g.Ww, w) : W3{0u, 0u, 0u};
    }
#pragma unroll
    for (int i = 0; i < NROWS; ++i) {
        const int dr = DR_LO + i, yy = y - dr;
        if (dr > dr_hi || yy < y_min) break;
        const int rp = g.rp[dr];
        if (dr > 0) {
            // cells that are not heads: responsible for the run of row yy whose head enters the window at
            // its right edge, column x + rp
            uint32_t resp = c & ~hd & (rp ? (ht[i].c >> rp) | (ht[i].p << (32 - rp)) : ht[i].c);
            while (resp) {
                const int x = 32 * w + __ffs(resp) - 1;
                resp &= resp - 1;
                run_pair<QUICK>(g, eps2, vxi, vyi, cimg, himg, par, y, x, yy, x + rp, -1, x + rp);
            }
        }
        // heads: every run inside the head's window
        uint32_t heads = hd;
        while (heads) {
            const int bit = __ffs(heads) - 1;
            heads &= heads - 1;
            const int x = 32 * w + bit;
            uint32_t cwin = win_bits(ct[i], bit, rp);
            const uint32_t hwin = win_bits(ht[i], bit, rp);
            if (dr == 0) cwin &= (1u << rp) - 1u;  // own row: the columns x - r .. x - 1
            // a run starts at every head bit and at the window's first core cell
            uint32_t starts = (hwin & cwin) | (cwin & (0u - cwin));
            while (starts) {
                const int j = __ffs(starts) - 1;
                starts &= starts - 1;
                const int xb = x - rp + j;
                run_pair<QUICK>(g, eps2, vxi, vyi, cimg, himg, par, y, x, yy, xb, x, ((hwin >> j) & 1u) ? xb : -1);
            }
        }
    }
}

// First union pass: a warp walks SWEEP_ROWS rows top-down (lane <-> word) and joins every run with the
// runs of its own row and of the row above.  Top-down order keeps the trees flat: what a run finds
// above it already hangs directly under its root, so a find is one or two hops, where joining all
// rows at once chains the runs of a tall region one under the other.  The row above a strip's first row
// is left to the second pass.
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_sweep(const float* __restrict__ vx, const float* __restrict__ vy,
                                                              const uint32_t* __restrict__ cbits,
                                                              const uint32_t* __restrict__ hbits, RunGeom g,
                                                              EpsTest eps2, int sweep_rows,
                                                              int32_t* __restrict__ parent) {
    const int strips = (g.H + sweep_rows - 1) / sweep_rows;
    const WarpPos p = warp_pos(g, strips);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const int w = p.wbase + p.lane;
    const int y0 = p.y * sweep_rows, y1 = min(y0 + sweep_rows, g.H);
    for (int y = y0; y < y1; ++y) {
        if (w < g.Ww)
            word_pairs<false, 0, 2>(g, eps2, vx + img * g.W, vy + img * g.W, cbits + img * g.Ww, hbits + img * g.Ww,
                                    parent + img * g.W, y, w, 1, y0);
        __syncwarp();
    }
}

// Second union pass (after a flatten): every pair of runs within reach, rows 0 .. floor(eps) above.
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_pairs(const float* __restrict__ vx, const float* __restrict__ vy,
                                                              const uint32_t* __restrict__ cbits,
                                                              const uint32_t* __restrict__ hbits, RunGeom g,
                                                              EpsTest eps2, int32_t* __restrict__ parent) {
    const WarpPos p = warp_pos(g, g.H);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const int w = p.wbase + p.lane;
    if (w >= g.Ww) return;
    const float* vxi = vx + img * g.W;
    const float* vyi = vy + img * g.W;
    const uint32_t* cimg = cbits + img * g.Ww;
    const uint32_t* himg = hbits + img * g.Ww;
    int32_t* par = parent + img * g.W;
    if (g.r <= 5) {   // the reference's eps: all six rows' words in flight at once
        word_pairs<true, 0, 6>(g, eps2, vxi, vyi, cimg, himg, par, p.y, w, g.r, 0);
    } else {
        word_pairs<true, 0, 4>(g, eps2, vxi, vyi, cimg, himg, par, p.y, w, 3, 0);
        word_pairs<true, 4, 4>(g, eps2, vxi, vyi, cimg, himg, par, p.y, w, 7, 0);
        word_pairs<true, 8, 4>(g, eps2, vxi, vyi, cimg, himg, par, p.y, w, 11, 0);
        word_pairs<true, 12, 4>(g, eps2, vxi, vyi, cimg, himg, par, p.y, w, min(g.r, 15), 0);
    }
}

// ---- flatten: heads point at their root; MARK: root bits + per-segment root counts ----------------
template <bool MARK>
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_flatten(const uint32_t* __restrict__ hbits, RunGeom g,
                                                                int32_t* __restrict__ parent,
                                                                uint32_t* __restrict__ rbits,
                                                                int32_t* __restrict__ seg_count) {
    const WarpPos p = warp_pos(g, g.H);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    int32_t* par = parent + img * g.W;
    const int w = p.wbase + p.lane;
    uint32_t h = w < g.Ww ? hbits[(img + p.y) * g.Ww + w] : 0u, roots = 0u;
    while (h) {
        const int bit = __ffs(h) - 1;
        h &= h - 1;
        const int a = p.y * g.W + 32 * w + bit;
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
        if (r == a) roots |= 1u << bit;
    }
    if (MARK) {
        if (w < g.Ww) rbits[(img + p.y) * g.Ww + w] = roots;
        const int cnt = __reduce_add_sync(0xffffffffu, __popc(roots));
        if (p.lane == 0) seg_count[(img + p.y) * g.nseg + p.seg] = cnt;
    }
}

// ---- labels ----------------------------------------------------------------------------------------
// lane <-> cell over the non-empty words of the row segment; the root lookup of the NEXT word is in
// flight while the current word's labels are fetched and stored.
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_labels(const float* __restrict__ vx, const float* __restrict__ vy,
                                                               const uint32_t* __restrict__ vbits,
                                                               const uint32_t* __restrict__ cbits,
                                                               const uint32_t* __restrict__ hbits, RunGeom g,
                                                               EpsTest eps2, const int32_t* __restrict__ parent,
                                                               const int32_t* __restrict__ rlabel,
                                                               const int32_t* __restrict__ seg_off, int cap,
                                                               int32_t* __restrict__ labels,
                                                               int32_t* __restrict__ indices) {
    const WarpPos p = warp_pos(g, g.H);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const uint32_t* cimg = cbits + img * g.Ww;
    const uint32_t* himg = hbits + img * g.Ww;
    const uint32_t* hrow = himg + static_cast<size_t>(p.y) * g.Ww;
    const float* vxi = vx + img * g.W;
    const float* vyi = vy + img * g.W;
    const int32_t* par = parent + img * g.W;
    const int32_t* rl = rlabel + img * g.W;
    const uint32_t mine = seg_word(vbits + (img + p.y) * g.Ww, g, p);
    unsigned nz = __ballot_sync(0xffffffffu, mine != 0u);
    if (nz == 0u) return;
    // row-major rank of the first valid cell of every word of the segment
    int before = __popc(mine);
#pragma unroll
    for (int o = 1; o < SEG_WORDS; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, before, o);
        if (p.lane >= o) before += v;
    }
    before += seg_off[(img + p.y) * g.nseg + p.seg] - __popc(mine);
    const uint32_t cmine = seg_word(cimg + static_cast<size_t>(p.y) * g.Ww, g, p);
    const uint32_t hmine = seg_word(hrow, g, p);
    // root of this lane's cell in word k when it is a core cell, else -1
    auto core_root = [&](int k) -> int {
        const uint32_t cword = __shfl_sync(0xffffffffu, cmine, k), hword = __shfl_sync(0xffffffffu, hmine, k);
        if (!((cword >> p.lane) & 1u)) return -1;
        const uint32_t m = hword & (0xffffffffu >> (31 - p.lane));  // the run's head, when it lies in this word
        const int a0 = m ? 32 * (p.wbase + k) + 31 - __clz(m) : head_of(hrow, 32 * (p.wbase + k) + p.lane);
        return par[p.y * g.W + a0];
    };
    int k = __ffs(nz) - 1;
    int root = core_root(k);
    while (true) {
        nz &= nz - 1;
        const int kn = nz ? __ffs(nz) - 1 : -1;
        const int root_n = kn >= 0 ? core_root(kn) : -1;
        const uint32_t word = __shfl_sync(0xffffffffu, mine, k);
        const int base = __shfl_sync(0xffffffffu, before, k);
        const int w = p.wbase + k, x = 32 * w + p.lane;
        const int slot = base + __popc(word & ((1u << p.lane) - 1u));
        if (((word >> p.lane) & 1u) && slot < cap) {
            if (root < 0) {
                // border cell: the smallest root among the core cells within eps
                const float vx0 = vxi[static_cast<size_t>(p.y) * g.W + x], vy0 = vyi[static_cast<size_t>(p.y) * g.W + x];
                for (int dr = -g.r; dr <= g.r; ++dr) {
                    const int yy = p.y + dr;
                    if (yy < 0 || yy >= g.H) continue;
                    const int rp = g.rp[dr < 0 ? -dr : dr];
                    uint32_t win = win_bits(load3(cimg + static_cast<size_t>(yy) * g.Ww, g.Ww, w), p.lane, rp);
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
        }
        if (kn < 0) break;
        k = kn, root = root_n;
    }
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
    for (int i = 0; i < 5; ++i) bump.take<uint32_t>(static_cast<size_t>(batch) * H * Ww);
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
    uint32_t* pbits = bump.take<uint32_t>(nw);
    uint32_t* hbits = bump.take<uint32_t>(nw);
    uint32_t* rbits = bump.take<uint32_t>(nw);
    int32_t* vseg = bump.take<int32_t>(ns);
    int32_t* rseg = bump.take<int32_t>(ns);
    int32_t* parent = bump.take<int32_t>(static_cast<size_t>(batch) * H * W);
    int32_t* rlabel = bump.take<int32_t>(static_cast<size_t>(batch) * H * W);
    int32_t* ncl = bump.take<int32_t>(batch);
    const int nblk = H * g.nseg, nt = 32 * RUN_WARPS;
    const dim3 grid(ceil_div(nblk, RUN_WARPS), batch);
    // rows one warp of the first union pass walks top-down
    static const int sweep_env = getenv("DATMO_SWEEP_ROWS") ? atoi(getenv("DATMO_SWEEP_ROWS")) : 0;
    const int sweep_rows = sweep_env > 0 ? sweep_env : 16;
    const dim3 grid_sweep(ceil_div(ceil_div(H, sweep_rows) * g.nseg, RUN_WARPS), batch);
    const int vec = (W & 15) == 0 && (reinterpret_cast<uintptr_t>(valid) & 15) == 0;
    cudaStream_t s = h->stream;
    {
        LaunchScope ls(h, tag(0));
        k_run_pack<<<grid, nt, 0, s>>>(valid, g, vec, vbits, vseg);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(0));
        k_run_scan<<<batch, 256, 0, s>>>(vseg, nblk, n_valid, nullptr, g, nullptr);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(1));
        k_run_core<<<grid, nt, 0, s>>>(vx_f, vy_f, vbits, g, eps2, cbits, pbits);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(2));
        k_run_heads<<<grid, nt, 0, s>>>(cbits, pbits, g, hbits, parent);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(5));
        k_run_sweep<<<grid_sweep, nt, 0, s>>>(vx_f, vy_f, cbits, hbits, g, eps2, sweep_rows, parent);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(3));
        k_run_flatten<false><<<grid, nt, 0, s>>>(hbits, g, parent, nullptr, nullptr);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(4));
        k_run_pairs<<<grid, nt, 0, s>>>(vx_f, vy_f, cbits, hbits, g, eps2, parent);
    }
    DATMO_POST_LAUNCH(h);
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
