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
//   pack      valid bytes -> valid bits (+ the per-segment counts the rank scan needs)
//   core      a valid cell is core when >= min_samples valid cells (itself included) satisfy
//             d2 = drow^2 + dcol^2 + dvx^2 + dvy^2 <= eps^2 (fp64, that order).  Inside a moving
//             region the two cells either side in the row and the cell above settle it from
//             shuffles; only the rest walks the window (rows nearest first, early exit).  The same
//             tests give the horizontal and vertical link bits of the valid cells.
//   heads     run = maximal chain of horizontally adjacent core cells that are pairwise within
//             eps; its head (first cell = minimum index) is the union-find node.  Every head
//             points at the run above its first vertical link — plain stores, no atomics; the
//             forest this builds already joins a convex region.
//   flatten   heads point at their root (pointer jumping; upper rows run first)
//   pairs     every pair of runs (A in row y, B in row y - dr) within reach of one another is
//             given to ONE cell of A — the first that sees B in its window: A's head for the
//             runs already in the head's window, else the cell at b0 - reach(dr).  Equal parents
//             prove "already joined" from two loads; otherwise the cell pairs of (A, B) are
//             tested until one is within eps.  Two or more rows apart, a run reached through an
//             unbroken column of vertical links is skipped without a look at the forest: the
//             row-1 pairs along that column make the union (6 000 runs, 13 000 run pairs that
//             reach the forest, 300 cell tests per 150 000-cell frame, against 9 million
//             candidate cell pairs).
//   flatten   again (over the head list), marking the roots (= minimum core index of each cluster)
//   ranks     label of a root = its rank among the roots = sklearn's cluster number
//   labels    core cell: label of its run's root; border cell: smallest root among the core
//             cells within eps (the cluster whose DFS reaches it first), else -1
#include <algorithm>
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
__global__ void __launch_bounds__(1024) k_run_scan(int32_t* __restrict__ seg_count, int nblk, int32_t* __restrict__ totals,
                                                  const uint32_t* __restrict__ rbits, RunGeom g,
                                                  int32_t* __restrict__ rlabel) {
    const int b = blockIdx.x;
    int32_t* s = seg_count + static_cast<size_t>(b) * nblk;
    __shared__ int s_part[32];
    const int nthr = blockDim.x, nwarp = nthr >> 5;   // a multiple of 32, at most 1024
    const int per = (nblk + nthr - 1) / nthr;
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
    for (int i = 0; i < nwarp; ++i) {
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

// ---- core cells + link bits of the valid cells -------------------------------------------------------
// cbits: core cells.  pbits / qbits: valid cells whose left / upper neighbour is valid and within eps
// (run links and vertical links once both cells turn out to be core).  A warp walks the non-empty
// words of its row segment with lane <-> cell; the velocities of the NEXT word are in flight while
// the current one is counted, the own row's window comes out of the words the warp already holds
// and the row neighbours' velocities out of shuffles.
__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_core(const float* __restrict__ vx, const float* __restrict__ vy,
                                                             const uint32_t* __restrict__ vbits, RunGeom g,
                                                             EpsTest eps2, uint32_t* __restrict__ cbits,
                                                             uint32_t* __restrict__ pbits,
                                                             uint32_t* __restrict__ qbits) {
    const WarpPos p = warp_pos(g, g.H);
    if (!p.live) return;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const uint32_t* vimg = vbits + img * g.Ww;
    const uint32_t* vrow = vimg + static_cast<size_t>(p.y) * g.Ww;
    const uint32_t* vup = vimg + static_cast<size_t>(max(p.y - 1, 0)) * g.Ww;
    const float* vxi = vx + img * g.W;
    const float* vyi = vy + img * g.W;
    const float* rx = vxi + static_cast<size_t>(p.y) * g.W;
    const float* ry = vyi + static_cast<size_t>(p.y) * g.W;
    const float* ux = rx - g.W;   // the row above (only read when p.y > 0)
    const float* uy = ry - g.W;
    const uint32_t mine = seg_word(vrow, g, p);
    const uint32_t above = p.y > 0 ? seg_word(vup, g, p) : 0u;
    uint32_t cmine = 0u, pmine = 0u, qmine = 0u;
    unsigned nz = __ballot_sync(0xffffffffu, mine != 0u);
    if (nz) {
        const uint32_t edge_l = p.wbase > 0 ? vrow[p.wbase - 1] : 0u;
        const uint32_t edge_r = p.wbase + SEG_WORDS < g.Ww ? vrow[p.wbase + SEG_WORDS] : 0u;
        struct Cell {
            uint32_t word, up;
            float vx, vy, uvx, uvy;
        };
        auto fetch = [&](int k) {
            Cell c;
            c.word = __shfl_sync(0xffffffffu, mine, k);
            c.up = __shfl_sync(0xffffffffu, above, k);
            c.vx = c.vy = c.uvx = c.uvy = 0.f;
            const int x = 32 * (p.wbase + k) + p.lane;
            if ((c.word >> p.lane) & 1u) {
                c.vx = rx[x], c.vy = ry[x];
                if ((c.up >> p.lane) & 1u) c.uvx = ux[x], c.uvy = uy[x];
            }
            return c;
        };
        int k = __ffs(nz) - 1;
        Cell cur = fetch(k);
        while (true) {
            nz &= nz - 1;
            const int kn = nz ? __ffs(nz) - 1 : -1;
            Cell nxt = cur;
            if (kn >= 0) nxt = fetch(kn);
            const int w = p.wbase + k, x = 32 * w + p.lane;
            W3 own;
            own.c = cur.word;
            own.m = __shfl_sync(0xffffffffu, mine, max(k - 1, 0));
            own.p = __shfl_sync(0xffffffffu, mine, min(k + 1, 31));
            if (k == 0) own.m = edge_l;
            if (k == 31) own.p = edge_r;
            const bool valid = (cur.word >> p.lane) & 1u;
            const float vx0 = cur.vx, vy0 = cur.vy;
            // the two cells either side in the row: velocities by shuffle (from memory at the word's ends)
            float nvx[4], nvy[4];   // columns x - 2, x - 1, x + 1, x + 2
            nvx[0] = __shfl_up_sync(0xffffffffu, vx0, 2), nvy[0] = __shfl_up_sync(0xffffffffu, vy0, 2);
            nvx[1] = __shfl_up_sync(0xffffffffu, vx0, 1), nvy[1] = __shfl_up_sync(0xffffffffu, vy0, 1);
            nvx[2] = __shfl_down_sync(0xffffffffu, vx0, 1), nvy[2] = __shfl_down_sync(0xffffffffu, vy0, 1);
            nvx[3] = __shfl_down_sync(0xffffffffu, vx0, 2), nvy[3] = __shfl_down_sync(0xffffffffu, vy0, 2);
            bool core = false, plink = false, qlink = false;
            if (valid) {
                const uint32_t near = win_bits(own, p.lane, 2);   // bits 0..4 <-> columns x - 2 .. x + 2
                const int dcs[4] = {-2, -1, 1, 2};
                int cnt = 1;   // the cell itself
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int dc = dcs[u];
                    if (!((near >> (dc + 2)) & 1u)) continue;
                    if (p.lane + dc < 0 || p.lane + dc > 31) nvx[u] = rx[x + dc], nvy[u] = ry[x + dc];
                    const bool in = within_eps(0, dc, vx0, vy0, nvx[u], nvy[u], eps2);
                    cnt += in;
                    if (dc == -1) plink = in;
                }
                if ((cur.up >> p.lane) & 1u) {
                    qlink = within_eps(1, 0, vx0, vy0, cur.uvx, cur.uvy, eps2);
                    cnt += qlink;
                }
                if (cnt < g.min_samples) {
                    // count over the whole window: rows in the order 0, -1, +1, -2, +2, .., early exit
                    cnt = 0;
                    for (int i = 0; i <= 2 * g.r && cnt < g.min_samples; ++i) {
                        const int dr = (i & 1) ? -((i + 1) >> 1) : (i >> 1);
                        const int yy = p.y + dr;
                        if (yy < 0 || yy >= g.H) continue;
                        const int rp = g.rp[dr < 0 ? -dr : dr];
                        uint32_t win =
                            win_bits(i == 0 ? own : load3(vimg + static_cast<size_t>(yy) * g.Ww, g.Ww, w), p.lane, rp);
                        const float* wx = vxi + static_cast<size_t>(yy) * g.W + (x - rp);
                        const float* wy = vyi + static_cast<size_t>(yy) * g.W + (x - rp);
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
                                if (j[u] >= 0) cx[u] = wx[j[u]], cy[u] = wy[j[u]];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (j[u] >= 0 && within_eps(dr, j[u] - rp, vx0, vy0, cx[u], cy[u], eps2)) ++cnt;
                        }
                    }
                }
                core = cnt >= g.min_samples;
            }
            const uint32_t cword = __ballot_sync(0xffffffffu, core), pword = __ballot_sync(0xffffffffu, plink),
                           qword = __ballot_sync(0xffffffffu, qlink);
            if (p.lane == k) cmine = cword, pmine = pword, qmine = qword;
            if (kn < 0) break;
            k = kn, cur = nxt;
        }
    }
    if (p.wbase + p.lane < g.Ww) {
        const size_t o = (img + p.y) * g.Ww + p.wbase + p.lane;
        cbits[o] = cmine, pbits[o] = pmine, qbits[o] = qmine;
    }
}

// ---- runs: head bits, vertical-link bits, first pointers (bit operations only; lane <-> word) ------
// head bits of a row from its core and horizontal-link words
__device__ __forceinline__ uint32_t head_word(const uint32_t* __restrict__ crow, const uint32_t* __restrict__ prow,
                                              int w) {
    const uint32_t c = crow[w];
    return c & ~(prow[w] & ((c << 1) | (w > 0 ? crow[w - 1] >> 31 : 0u)));
}
// cells bit .. (first stop above bit) - 1 of a word: the part of the run headed at `bit` that lies in it
__device__ __forceinline__ uint32_t run_part(uint32_t c, uint32_t h, int bit) {
    const uint32_t from = 0xffffffffu << bit;
    const uint32_t stop = (~c | h) & (from << 1);
    return stop ? from & ((1u << (__ffs(stop) - 1)) - 1u) : from;
}

__global__ void __launch_bounds__(32 * RUN_WARPS) k_run_heads(const uint32_t* __restrict__ cbits,
                                                              const uint32_t* __restrict__ pbits,
                                                              const uint32_t* __restrict__ qbits, RunGeom g,
                                                              uint32_t* __restrict__ hbits,
                                                              uint32_t* __restrict__ ubits,
                                                              int32_t* __restrict__ parent,
                                                              int32_t* __restrict__ hlist, int hcap,
                                                              int32_t* __restrict__ hcount,
                                                              uint32_t* __restrict__ rbits,
                                                              int32_t* __restrict__ rseg) {
    const WarpPos p = warp_pos(g, g.H);
    if (!p.live) return;
    const int w = p.wbase + p.lane;
    const bool on = w < g.Ww;
    const size_t img = static_cast<size_t>(p.b) * g.H;
    const uint32_t* crow = cbits + (img + p.y) * g.Ww;
    const uint32_t* prow = pbits + (img + p.y) * g.Ww;
    const uint32_t c = on ? crow[w] : 0u;
    const uint32_t h = on ? head_word(crow, prow, w) : 0u;
    // vertical links: both cells core and within eps
    const uint32_t u = on && p.y > 0 ? qbits[(img + p.y) * g.Ww + w] & c & (crow - g.Ww)[w] : 0u;
    if (on) {
        hbits[(img + p.y) * g.Ww + w] = h;
        ubits[(img + p.y) * g.Ww + w] = u;
        rbits[(img + p.y) * g.Ww + w] = 0u;   // the root plane the last flatten marks (saves two memset launches)
    }
    if (p.lane == 0) rseg[(img + p.y) * g.nseg + p.seg] = 0;
    // the frame's list of heads (any order): one atomic per warp reserves the slots
    int slot = __popc(h);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, slot, o);
        if (p.lane >= o) slot += v;
    }
    const int total = __shfl_sync(0xffffffffu, slot, 31);
    if (total == 0) return;
    int base = 0;
    if (p.lane == 0) base = atomicAdd(hcount + p.b, total);
    slot += __shfl_sync(0xffffffffu, base, 0) - __popc(h);
    int32_t* list = hlist + static_cast<size_t>(p.b) * hcap;
    int32_t* par = parent + img * g.W;
    uint32_t hh = h;
    while (hh) {
        const int bit = __ffs(hh) - 1;
        hh &= hh - 1;
        const int a = p.y * g.W + 32 * w + bit;
        if (slot < hcap) list[slot] = a;
        ++slot;
        int target = a;
        const uint32_t up = u & run_part(c, h, bit);
        if (up) {
            // the run of the row above that holds the column of the first vertical link; its head comes
            // out of that row's core / link words (its head bits are being written by another warp)
            const int xc = 32 * w + __ffs(up) - 1;
            int wa = xc >> 5;
            uint32_t m = head_word(crow - g.Ww, prow - g.Ww, wa) & (0xffffffffu >> (31 - (xc & 31)));
            while (m == 0u && wa > 0) m = head_word(crow - g.Ww, prow - g.Ww, --wa);
            target = (p.y - 1) * g.W + 32 * wa + 31 - __clz(m);
        }
        par[a] = target;
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

// A pair of runs whose first cell pair failed the eps test: the remaining cell pairs are walked by the
// whole warp (scan_pair), not by the one thread that owns the pair.
struct PendingScan {
    int y, x, yy, b0, a1, b1, ra, rb;   // y < 0: nothing pending
};

// Run of (y, x) against the run of row yy = y - dr whose first cell inside the window is column xb
// and whose head is b0; (y, x) is the cell responsible for the pair, a0 its run's head.  The forest
// was flattened before this kernel, so equal parents prove "already joined" from two independent
// loads.  Otherwise the first cell pair (x, xb) is tested; when that fails too, the pair is left in
// `pend` for the warp (or, if `pend` is taken, walked here).
__device__ __noinline__ void run_pair(const RunGeom& g, const EpsTest& eps2, const float* __restrict__ vxi,
                                      const float* __restrict__ vyi, const uint32_t* __restrict__ cimg,
                                      const uint32_t* __restrict__ himg, int32_t* par, int y, int x, int yy, int xb,
                                      int a0, int b0, PendingScan& pend) {
    int ra = y * g.W + a0, rb = yy * g.W + b0;
    const int pa = __ldcg(par + ra), pb = __ldcg(par + rb);
    if (pa == pb) return;
    const float* ax = vxi + static_cast<size_t>(y) * g.W;
    const float* ay = vyi + static_cast<size_t>(y) * g.W;
    const float* bx = vxi + static_cast<size_t>(yy) * g.W;
    const float* by = vyi + static_cast<size_t>(yy) * g.W;
    const float vxa = ax[x], vya = ay[x], vxb = bx[xb], vyb = by[xb];
    if (pa != ra) ra = uf_find(par, pa);
    if (pb != rb) rb = uf_find(par, pb);
    if (ra == rb) return;
    const int dr = y - yy, rp = g.rp[dr];
    if (within_eps(dr, x - xb, vxa, vya, vxb, vyb, eps2)) {
        uf_link(par, ra, rb);
        return;
    }
    const uint32_t* hrow_a = himg + static_cast<size_t>(y) * g.Ww;
    const uint32_t* hrow_b = himg + static_cast<size_t>(yy) * g.Ww;
    const int a1 = run_end(cimg + static_cast<size_t>(y) * g.Ww, hrow_a, g.Ww, x);
    const int b1 = run_end(cimg + static_cast<size_t>(yy) * g.Ww, hrow_b, g.Ww, b0);
    if (pend.y < 0) {
        pend = PendingScan{y, x, yy, b0, a1, b1, ra, rb};
        return;
    }
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

// All 32 lanes walk the cell pairs of every pending pair of the warp, one pair at a time: lane l takes
// the cells x + l, x + l + 32, .. of run A against their windows in run B; the first hit joins the runs.
__device__ __forceinline__ void scan_pending(const RunGeom& g, const EpsTest& eps2, const float* __restrict__ vxi,
                                             const float* __restrict__ vyi, int32_t* par, PendingScan& pend) {
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(0xffffffffu, pend.y >= 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int y = __shfl_sync(0xffffffffu, pend.y, src), x = __shfl_sync(0xffffffffu, pend.x, src);
        const int yy = __shfl_sync(0xffffffffu, pend.yy, src), b0 = __shfl_sync(0xffffffffu, pend.b0, src);
        const int a1 = __shfl_sync(0xffffffffu, pend.a1, src), b1 = __shfl_sync(0xffffffffu, pend.b1, src);
        const int dr = y - yy, rp = g.rp[dr];
        const float* ax = vxi + static_cast<size_t>(y) * g.W;
        const float* ay = vyi + static_cast<size_t>(y) * g.W;
        const float* bx = vxi + static_cast<size_t>(yy) * g.W;
        const float* by = vyi + static_cast<size_t>(yy) * g.W;
        const int last = min(a1, b1 + rp);   // cells of A beyond it do not see B
        bool hit = false;
        for (int base = x; base <= last && !hit; base += 32) {
            const int xa = base + lane;
            bool mine = false;
            if (xa <= last) {
                const int lo = max(b0, xa - rp), hi = dr > 0 ? min(b1, xa + rp) : min(b1, xa - 1);
                const float vx0 = ax[xa], vy0 = ay[xa];
                for (int q = lo; q <= hi && !mine; ++q) mine = within_eps(dr, xa - q, vx0, vy0, bx[q], by[q], eps2);
            }
            hit = __any_sync(0xffffffffu, mine);
        }
        if (hit && lane == src) uf_link(par, pend.ra, pend.rb);
    }
    pend.y = -1;
}

// Union pass (after the first pointers were flattened), one thread per run head: every pair of runs
// within reach of one another is seen by exactly one head.
//   looking up    (rows y - dr, dr = 0 .. floor(eps); dr = 0: the columns left of the head) every run
//                 with a cell inside the head's window.  Two or more rows up, the run an unbroken column
//                 of vertical links leads to is skipped: the row-1 pairs along that column join it.
//   looking down  (rows y + dr) the one run that covers the window's left edge from further left: its
//                 own head does not see this run, and its first cell that does is x - reach(dr).
// a pair of runs a head is responsible for: the cell (ya, xa) of run A (head a0) against the run of row
// yb whose first cell in reach is xb (head b0)
struct RunCand {
    int ya, xa, yb, xb, a0, b0;
};
constexpr int PAIR_ROWS = 3;    // rows whose words are fetched together
constexpr int PAIR_QUEUE = 6;   // pairs a head collects before any touches the forest

__global__ void __launch_bounds__(128) k_run_pairs(const float* __restrict__ vx, const float* __restrict__ vy,
                                                   const uint32_t* __restrict__ cbits,
                                                   const uint32_t* __restrict__ hbits,
                                                   const uint32_t* __restrict__ ubits, RunGeom g, EpsTest eps2,
                                                   const int32_t* __restrict__ hlist, int hcap,
                                                   const int32_t* __restrict__ hcount, int dr_min, int dr_max,
                                                   int32_t* __restrict__ parent) {
    const int b = blockIdx.y;
    const int n = min(hcount[b], hcap);
    const size_t img = static_cast<size_t>(b) * g.H;
    const float* vxi = vx + img * g.W;
    const float* vyi = vy + img * g.W;
    const uint32_t* cimg = cbits + img * g.Ww;
    const uint32_t* himg = hbits + img * g.Ww;
    const uint32_t* uimg = ubits + img * g.Ww;
    int32_t* par = parent + img * g.W;
    const int32_t* list = hlist + static_cast<size_t>(b) * hcap;
    const int lane = threadIdx.x & 31;
    PendingScan pend;
    pend.y = -1;
    // warp-uniform trip count: the lanes past the end of the list still take part in the shared scans
    for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        RunCand q[PAIR_QUEUE];
        int nq = 0;
        if (i < n) {
            // 1. the head's pairs, from bit operations on the words of the rows around it.  The words of
            //    PAIR_ROWS rows (up and down) are fetched before any is looked at; a pair waits in q so
            //    that all lanes of the warp walk the forest together afterwards.
            const int cell = list[i];
            const int y = cell / g.W, x = cell - y * g.W, w = x >> 5, bit = x & 31;
            const uint32_t c = cimg[static_cast<size_t>(y) * g.Ww + w], hd = himg[static_cast<size_t>(y) * g.Ww + w];
            const uint32_t part = run_part(c, hd, bit);
            uint32_t chain = 0xffffffffu;   // AND of the vertical-link words of rows y .. y - dr + 1
            auto push = [&](int ya, int xa, int yb, int xb, int a0, int b0) {
                if (nq < PAIR_QUEUE)
                    q[nq++] = RunCand{ya, xa, yb, xb, a0, b0};
                else
                    run_pair(g, eps2, vxi, vyi, cimg, himg, par, ya, xa, yb, xb, a0, b0, pend);
            };
            if (dr_min > 0) chain = uimg[static_cast<size_t>(y) * g.Ww + w];   // rows y .. y - dr_min + 1 (dr_min <= 1)
            for (int d0 = dr_min; d0 <= dr_max; d0 += PAIR_ROWS) {
                W3 ht[PAIR_ROWS], ct[PAIR_ROWS];
                uint32_t ut[PAIR_ROWS], cd[PAIR_ROWS], hdn[PAIR_ROWS];
#pragma unroll
                for (int k = 0; k < PAIR_ROWS; ++k) {
                    const int dr = d0 + k, yy = y - dr, yd = y + dr;
                    const bool up = dr <= dr_max && yy >= 0;
                    ht[k] = up ? load3(himg + static_cast<size_t>(yy) * g.Ww, g.Ww, w) : W3{0u, 0u, 0u};
                    ct[k] = up ? load3(cimg + static_cast<size_t>(yy) * g.Ww, g.Ww, w) : W3{0u, 0u, 0u};
                    ut[k] = up ? uimg[static_cast<size_t>(yy) * g.Ww + w] : 0u;
                    const int xl = x - g.rp[min(dr, g.r)];
                    const bool down = dr > 0 && dr <= dr_max && yd < g.H && xl >= 0;
                    cd[k] = down ? cimg[static_cast<size_t>(yd) * g.Ww + (xl >> 5)] : 0u;
                    hdn[k] = down ? himg[static_cast<size_t>(yd) * g.Ww + (xl >> 5)] : 0u;
                }
#pragma unroll
                for (int k = 0; k < PAIR_ROWS; ++k) {
                    const int dr = d0 + k;
                    if (dr > dr_max) break;
                    const int rp = g.rp[dr];
                    // looking down: the run of row y + dr that covers the window's left edge from further left
                    const int yd = y + dr, xl = x - rp;
                    if (xl >= 0 && (((cd[k] & ~hdn[k]) >> (xl & 31)) & 1u))
                        push(yd, xl, y, x, head_of(himg + static_cast<size_t>(yd) * g.Ww, xl), x);
                    // looking up (dr = 0: the columns left of the head): every run inside the window
                    const int yy = y - dr;
                    if (yy < 0) continue;
                    const uint32_t* hrow_b = himg + static_cast<size_t>(yy) * g.Ww;
                    uint32_t cwin = win_bits(ct[k], bit, rp);
                    const uint32_t hwin = win_bits(ht[k], bit, rp);
                    if (dr == 0) cwin &= (1u << rp) - 1u;
                    // a run starts at every head bit and at the window's first core cell
                    uint32_t starts = (hwin & cwin) | (cwin & (0u - cwin));
                    if (starts) {
                        // two or more rows up: the run an unbroken column of vertical links leads to is
                        // joined by the row-1 pairs along that column
                        int implied = -1;
                        if (dr >= 2) {
                            const uint32_t col = chain & part;
                            if (col) implied = head_of(hrow_b, 32 * w + __ffs(col) - 1);
                        }
                        while (starts) {
                            const int j = __ffs(starts) - 1;
                            starts &= starts - 1;
                            const int xb = x - rp + j;
                            const int b0 = ((hwin >> j) & 1u) ? xb : head_of(hrow_b, xb);
                            if (b0 != implied) push(y, x, yy, xb, x, b0);
                        }
                    }
                    chain &= ut[k];   // rows y .. yy: what a pair dr + 1 rows apart may rely on
                }
            }
        }
        // 2. the forest: every lane's j-th pair at the same time
        for (int j = 0; j < PAIR_QUEUE; ++j) {
            if (__ballot_sync(0xffffffffu, j < nq) == 0u) break;
            if (j < nq)
                run_pair(g, eps2, vxi, vyi, cimg, himg, par, q[j].ya, q[j].xa, q[j].yb, q[j].xb, q[j].a0, q[j].b0, pend);
        }
        // 3. pairs whose first cells were not within eps: the warp walks the rest together
        scan_pending(g, eps2, vxi, vyi, par, pend);
    }
}

// ---- flatten over the head list (between the union passes): one thread per head ---------------------
// rbits / seg_count (both zeroed by the caller) given: the roots are marked and counted per segment as well
__global__ void __launch_bounds__(128) k_run_flatten_list(const int32_t* __restrict__ hlist, int hcap,
                                                          const int32_t* __restrict__ hcount, RunGeom g,
                                                          int32_t* __restrict__ parent, uint32_t* __restrict__ rbits,
                                                          int32_t* __restrict__ seg_count) {
    const int b = blockIdx.y;
    const int n = min(hcount[b], hcap);
    const size_t img = static_cast<size_t>(b) * g.H;
    int32_t* par = parent + img * g.W;
    const int32_t* list = hlist + static_cast<size_t>(b) * hcap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int a = list[i];
        // every link is finished (previous kernel): roots are fixed points, and concurrent compressions
        // only replace a parent by one of its ancestors.  Progress is published hop by hop.
        int r = __ldcg(par + a);
        while (true) {
            const int up = __ldcg(par + r);
            if (up == r) break;
            __stcg(par + a, up);
            r = up;
        }
        if (rbits != nullptr && r == a) {   // ~150 roots per frame
            const int y = a / g.W, x = a - y * g.W;
            atomicOr(rbits + (img + y) * g.Ww + (x >> 5), 1u << (x & 31));
            atomicAdd(seg_count + (img + y) * g.nseg + (x >> 5) / SEG_WORDS, 1);
        }
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
    for (int i = 0; i < 7; ++i) bump.take<uint32_t>(static_cast<size_t>(batch) * H * Ww);
    for (int i = 0; i < 2; ++i) bump.take<int32_t>(static_cast<size_t>(batch) * H * nseg);
    for (int i = 0; i < 2; ++i) bump.take<int32_t>(static_cast<size_t>(batch) * H * W);
    bump.take<int32_t>(static_cast<size_t>(batch) * (static_cast<size_t>(H) * ((W + 1) / 2)));
    for (int i = 0; i < 2; ++i) bump.take<int32_t>(batch);
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
    uint32_t* vbits = bump.take<uint32_t>(nw);   // valid cells
    uint32_t* cbits = bump.take<uint32_t>(nw);   // core cells
    uint32_t* pbits = bump.take<uint32_t>(nw);   // valid, left neighbour valid and within eps
    uint32_t* qbits = bump.take<uint32_t>(nw);   // valid, upper neighbour valid and within eps
    uint32_t* hbits = bump.take<uint32_t>(nw);   // run heads
    uint32_t* ubits = bump.take<uint32_t>(nw);   // vertical links between core cells
    uint32_t* rbits = bump.take<uint32_t>(nw);   // roots
    int32_t* vseg = bump.take<int32_t>(ns);
    int32_t* rseg = bump.take<int32_t>(ns);
    int32_t* parent = bump.take<int32_t>(static_cast<size_t>(batch) * H * W);   // written at run heads only
    int32_t* rlabel = bump.take<int32_t>(static_cast<size_t>(batch) * H * W);   // written at roots only
    const int hcap = H * ((W + 1) / 2);   // a row holds at most ceil(W / 2) run heads
    int32_t* hlist = bump.take<int32_t>(static_cast<size_t>(batch) * hcap);   // run heads of every frame, any order
    int32_t* ncl = bump.take<int32_t>(batch);
    int32_t* hcount = bump.take<int32_t>(batch);
    DATMO_CHECK_CUDA(h, cudaMemsetAsync(hcount, 0, sizeof(int32_t) * batch, h->stream));
    const int nblk = H * g.nseg, nt = 32 * RUN_WARPS;
    // one CTA per frame scans its segment counts: as many threads as segments, up to 1024 (short serial chains)
    const int scan_threads = std::min(1024, std::max(256, (nblk + 31) & ~31));
    const dim3 grid(ceil_div(nblk, RUN_WARPS), batch);
    const int vec = (W & 15) == 0 && (reinterpret_cast<uintptr_t>(valid) & 15) == 0;
    cudaStream_t s = h->stream;
    {
        LaunchScope ls(h, tag(0));
        k_run_pack<<<grid, nt, 0, s>>>(valid, g, vec, vbits, vseg);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(0));
        k_run_scan<<<batch, scan_threads, 0, s>>>(vseg, nblk, n_valid, nullptr, g, nullptr);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(1));
        k_run_core<<<grid, nt, 0, s>>>(vx_f, vy_f, vbits, g, eps2, cbits, pbits, qbits);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(2));
        k_run_heads<<<grid, nt, 0, s>>>(cbits, pbits, qbits, g, hbits, ubits, parent, hlist, hcap, hcount, rbits, rseg);
    }
    DATMO_POST_LAUNCH(h);
    // one thread per run head, grid-stride over the frame's list (its length lives on the device)
    const dim3 grid_list(std::max(16, std::min(64, ceil_div(16 * h->sm_count, batch))), batch);
    {
        LaunchScope ls(h, tag(3));
        k_run_flatten_list<<<grid_list, 128, 0, s>>>(hlist, hcap, hcount, g, parent, nullptr, nullptr);
    }
    DATMO_POST_LAUNCH(h);
    // runs side by side in one row first: regions that touch sideways become one tree before the pass over
    // all rows, which then finds almost every pair joined (8 845 -> 1 674 pairs that need the forest on a
    // 260 000-cell frame)
    {
        LaunchScope ls(h, tag(4));
        k_run_pairs<<<grid_list, 128, 0, s>>>(vx_f, vy_f, cbits, hbits, ubits, g, eps2, hlist, hcap, hcount, 0, 0, parent);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(3));
        k_run_flatten_list<<<grid_list, 128, 0, s>>>(hlist, hcap, hcount, g, parent, nullptr, nullptr);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(4));
        k_run_pairs<<<grid_list, 128, 0, s>>>(vx_f, vy_f, cbits, hbits, ubits, g, eps2, hlist, hcap, hcount, 1, g.r, parent);
    }
    DATMO_POST_LAUNCH(h);
    // last flatten over the head list; it marks and counts the roots in the root plane (zeroed by k_run_heads)
    {
        LaunchScope ls(h, tag(3));
        k_run_flatten_list<<<grid_list, 128, 0, s>>>(hlist, hcap, hcount, g, parent, rbits, rseg);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, tag(0));
        k_run_scan<<<batch, scan_threads, 0, s>>>(rseg, nblk, n_clusters ? n_clusters : ncl, rbits, g, rlabel);
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
