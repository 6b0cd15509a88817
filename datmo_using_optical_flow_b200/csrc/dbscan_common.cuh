// Pieces shared by the two DBSCAN implementations (dbscan.cu: cell-level passes, any eps;
// dbscan_runs.cu: row runs over bit-packed grids, the fast path): the exact neighbour
// predicate of sklearn's radius query and the min-root union-find.  sm_100a.
#pragma once

#include "common.cuh"

namespace {

// ---- neighbour predicate ----------------------------------------------------------------
// eps^2 with an f32 guard band: the f32 evaluation of d2 is within 3e-7 (relative) of the fp64 one,
// so outside [lo, hi] = eps^2 (1 -+ 2e-6) it decides; only the sliver in between pays for fp64.
struct EpsTest {
    double e2;
    float lo, hi;
};

__device__ __forceinline__ bool within_eps(int dr, int dc, float vx0, float vy0, float vx1, float vy1,
                                           const EpsTest& e) {
    const float fvx = vx0 - vx1, fvy = vy0 - vy1;
    const float s = static_cast<float>(dr * dr + dc * dc) + fvx * fvx + fvy * fvy;
    if (s < e.lo) return true;
    if (s > e.hi) return false;
    double d2 = static_cast<double>(dr * dr);
    d2 = __dadd_rn(d2, static_cast<double>(dc * dc));
    const double dvx = __dsub_rn(static_cast<double>(vx0), static_cast<double>(vx1));
    const double dvy = __dsub_rn(static_cast<double>(vy0), static_cast<double>(vy1));
    d2 = __dadd_rn(d2, __dmul_rn(dvx, dvx));
    d2 = __dadd_rn(d2, __dmul_rn(dvy, dvy));
    return d2 <= e.e2;
}

// Union-find over int32 cell indices.  Other CTAs link roots concurrently, and L1 is not
// coherent between SMs, so every read of the forest goes to L2 (__ldcg) and every link is
// an atomicCAS; a stale L1 line could otherwise make a thread retry the same CAS forever.
__device__ __forceinline__ int uf_find(int32_t* parent, int a) {
    int p = __ldcg(parent + a);
    while (p != a) {
        int gp = __ldcg(parent + p);
        if (gp != p) __stcg(parent + a, gp);  // path halving; only ever shortens the path
        a = p;
        p = gp;
    }
    return a;
}

__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) {
            int t = a;
            a = b;
            b = t;
        }
        // a > b: hang the larger root under the smaller, so a root is always the
        // minimum index of its component
        int old = atomicCAS(parent + a, a, b);
        if (old == a) return;
        a = old;  // someone linked a first; continue from where it points now
    }
}

}  // namespace
