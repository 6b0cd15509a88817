// RANSAC ground-plane segmentation, sm_100a.
//
// Replaces flipped_pcd.segment_plane(distance_threshold=0.5, ransac_n=5,
// num_iterations=5000) at Optical_flow/main.py:73 (Open3D; absent from this image, so
// the algorithm is the published one as restated in oracle/ransac_np.py — PARITY
// UNPINNED against the real library).
//
// All hypotheses are independent, so they are scored as one batch instead of Open3D's
// iteration loop:
//   k_ransac_hyp    one thread per hypothesis: counter-hash sample indices, plane through
//                   the samples (3 points: triangle normal; more: centroid + covariance
//                   cofactors), fp64
//   k_ransac_score  a CTA holds a tile of points in shared memory and 256 hypotheses in
//                   registers, one per thread; every thread walks the tile (broadcast
//                   shared-memory reads) testing |a x + b y + c z + d| < thr in fp64 and keeps
//                   its own inlier count / error sum: no atomics, no cross-thread traffic
//                   in the hot loop; per-tile partials are reduced in tile order
//                   (deterministic)
//   k_ransac_upper / _lo_seed   f32 upper bound of every hypothesis' count (packed FFMA2, 4 tests per three
//                   16-byte broadcast loads) and the lower bound of the most promising one: only the
//                   hypotheses that can still win are scored in fp64 (same winner as scoring all of them)
//   k_ransac_best   warp-shuffle arg-max over (count desc, rmse asc, index asc)
//   k_ransac_mask   inlier mask of the winner + moments for the refit
#include <math.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int RS_ATTEMPTS = 16;
constexpr int RS_MAX_N = 8;        // samples per hypothesis
constexpr int RS_TILE = 2048;      // points per CTA tile
constexpr int RS_THREADS = 256;    // hypotheses per CTA

__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int LAYOUT>
__device__ __forceinline__ void load_point(const void* pts, int64_t i, int flip_x, double& x, double& y, double& z) {
    if (LAYOUT == DATMO_PTS_F64_XYZ) {
        const double* p = static_cast<const double*>(pts) + 3 * i;
        x = p[0], y = p[1], z = p[2];
    } else {
        const float4 p = static_cast<const float4*>(pts)[i];
        x = p.x, y = p.y, z = p.z;
    }
    if (flip_x) x = -x;
}

// GetPlaneFromPoints on moments: centroid c, deviations' second moments
__device__ __forceinline__ void plane_from_moments(double cx, double cy, double cz, double xx, double xy, double xz,
                                                   double yy, double yz, double zz, double out[4]) {
    const double det_x = yy * zz - yz * yz;
    const double det_y = xx * zz - xz * xz;
    const double det_z = xx * yy - xy * xy;
    double a, b, c;
    if (det_x > det_y && det_x > det_z) {
        a = det_x, b = xz * yz - xy * zz, c = xy * yz - xz * yy;
    } else if (det_y > det_z) {
        a = xz * yz - xy * zz, b = det_y, c = xy * xz - yz * xx;
    } else {
        a = xy * yz - xz * yy, b = xy * xz - yz * xx, c = det_z;
    }
    const double norm = sqrt(a * a + b * b + c * c);
    if (!(norm > 0.0) || !isfinite(norm)) {
        out[0] = out[1] = out[2] = out[3] = 0.0;
        return;
    }
    a /= norm, b /= norm, c /= norm;
    out[0] = a, out[1] = b, out[2] = c;
    out[3] = -(a * cx + b * cy + c * cz);
}

template <int LAYOUT>
__global__ void __launch_bounds__(128) k_ransac_hyp(const void* __restrict__ pts, int64_t n, int flip_x, int ransac_n,
                                                    int iters, uint64_t seed, double* __restrict__ planes) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= iters) return;
    int64_t idx[RS_MAX_N];
    int filled = 0;
    for (int t = 0; t < RS_ATTEMPTS && filled < ransac_n; ++t) {
        int64_t cand = static_cast<int64_t>(mix64(seed, static_cast<uint64_t>(it) * RS_ATTEMPTS + t) %
                                            static_cast<uint64_t>(n));
        bool dup = false;
        for (int j = 0; j < filled; ++j) dup |= idx[j] == cand;
        if (!dup) idx[filled++] = cand;
    }
    double out[4] = {0, 0, 0, 0};
    if (filled >= ransac_n) {
        double px[RS_MAX_N], py[RS_MAX_N], pz[RS_MAX_N];
        for (int j = 0; j < ransac_n; ++j) load_point<LAYOUT>(pts, idx[j], flip_x, px[j], py[j], pz[j]);
        if (ransac_n == 3) {
            const double e0x = px[1] - px[0], e0y = py[1] - py[0], e0z = pz[1] - pz[0];
            const double e1x = px[2] - px[0], e1y = py[2] - py[0], e1z = pz[2] - pz[0];
            double a = e0y * e1z - e0z * e1y, b = e0z * e1x - e0x * e1z, c = e0x * e1y - e0y * e1x;
            const double norm = sqrt(a * a + b * b + c * c);
            if (norm > 0.0 && isfinite(norm)) {
                a /= norm, b /= norm, c /= norm;
                out[0] = a, out[1] = b, out[2] = c;
                out[3] = -(a * px[0] + b * py[0] + c * pz[0]);
            }
        } else {
            double cx = 0, cy = 0, cz = 0;
            for (int j = 0; j < ransac_n; ++j) cx += px[j], cy += py[j], cz += pz[j];
            cx /= ransac_n, cy /= ransac_n, cz /= ransac_n;
            double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
            for (int j = 0; j < ransac_n; ++j) {
                const double rx = px[j] - cx, ry = py[j] - cy, rz = pz[j] - cz;
                xx += rx * rx, xy += rx * ry, xz += rx * rz, yy += ry * ry, yz += ry * rz, zz += rz * rz;
            }
            plane_from_moments(cx, cy, cz, xx, xy, xz, yy, yz, zz, out);
        }
    }
    double* o = planes + 4 * static_cast<size_t>(it);
    o[0] = out[0], o[1] = out[1], o[2] = out[2], o[3] = out[3];
}

__device__ __forceinline__ double plane_dist(double a, double b, double c, double d, double x, double y, double z) {
    // |((a*x + b*y) + c*z) + d|, every operation rounded (matches the oracle bit for bit)
    return fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, x), __dmul_rn(b, y)), __dmul_rn(c, z)), d));
}

// grid: (point tiles, hypothesis groups).  part_*: [n_tiles][iters]
// The error sum of a (tile, hypothesis) is defined chunk-wise — 32 chunks of RS_TILE / 32 consecutive points,
// each summed in point order, the chunk sums added in chunk order — so that one thread walking the tile (all
// hypotheses) and a warp with a lane per chunk (list mode) produce the same bits.
// list / n_list: the exact pass over the candidates the f32 bounds left (k_ransac_upper, k_ransac_lo_seed):
// grid (point tiles, 1), every warp takes the slots warp, warp + 8, .. < *n_list, a lane per chunk; the
// partials are indexed by slot.
constexpr int RS_CHUNK = RS_TILE / 32;

template <int LAYOUT>
__global__ void __launch_bounds__(RS_THREADS) k_ransac_score(const void* __restrict__ pts, int64_t n, int flip_x,
                                                             const double* __restrict__ planes, int iters, double thr,
                                                             int32_t* __restrict__ part_cnt,
                                                             double* __restrict__ part_err,
                                                             const int32_t* __restrict__ list,
                                                             const int32_t* __restrict__ n_list) {
    __shared__ double sx[RS_TILE], sy[RS_TILE], sz[RS_TILE];
    const int n_slots = list ? min(*n_list, iters) : iters;
    if (n_slots == 0 || (!list && static_cast<int>(blockIdx.y) * RS_THREADS >= n_slots)) return;
    const int64_t p0 = static_cast<int64_t>(blockIdx.x) * RS_TILE;
    const int np = static_cast<int>(min(static_cast<int64_t>(RS_TILE), n - p0));
    for (int i = threadIdx.x; i < np; i += RS_THREADS) load_point<LAYOUT>(pts, p0 + i, flip_x, sx[i], sy[i], sz[i]);
    __syncthreads();
    if (list) {
        const int lane = threadIdx.x & 31;
        const int i0 = lane * RS_CHUNK, i1 = min(i0 + RS_CHUNK, np);
        for (int slot = threadIdx.x >> 5; slot < n_slots; slot += RS_THREADS / 32) {
            const int hyp = list[slot];
            const double a = planes[4 * hyp], b = planes[4 * hyp + 1], c = planes[4 * hyp + 2], d = planes[4 * hyp + 3];
            int cnt = 0;
            double s = 0.0;
            if (a != 0.0 || b != 0.0 || c != 0.0 || d != 0.0) {
#pragma unroll 4
                for (int i = i0; i < i1; ++i) {
                    const double dist = plane_dist(a, b, c, d, sx[i], sy[i], sz[i]);
                    if (dist < thr) {
                        ++cnt;
                        s += dist;
                    }
                }
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            double err = 0.0;
            for (int l = 0; l < 32; ++l) err += __shfl_sync(0xffffffffu, s, l);   // chunk order
            if (lane == 0) {
                part_cnt[static_cast<size_t>(blockIdx.x) * iters + slot] = cnt;
                part_err[static_cast<size_t>(blockIdx.x) * iters + slot] = err;
            }
        }
        return;
    }
    const int slot = blockIdx.y * RS_THREADS + threadIdx.x;
    if (slot >= n_slots) return;
    const int hyp = slot;
    const double a = planes[4 * hyp], b = planes[4 * hyp + 1], c = planes[4 * hyp + 2], d = planes[4 * hyp + 3];
    int cnt = 0;
    double err = 0.0;
    if (a != 0.0 || b != 0.0 || c != 0.0 || d != 0.0) {
        for (int i0 = 0; i0 < np; i0 += RS_CHUNK) {
            const int i1 = min(i0 + RS_CHUNK, np);
            double s = 0.0;
#pragma unroll 4
            for (int i = i0; i < i1; ++i) {
                const double dist = plane_dist(a, b, c, d, sx[i], sy[i], sz[i]);
                if (dist < thr) {
                    ++cnt;
                    s += dist;
                }
            }
            err += s;
        }
    }
    part_cnt[static_cast<size_t>(blockIdx.x) * iters + slot] = cnt;
    part_err[static_cast<size_t>(blockIdx.x) * iters + slot] = err;
}

// ---- f32 bounds on the hypotheses' inlier counts ---------------------------------------------------
// The exact score needs an fp64 distance per (hypothesis, point) — 1.2e9 of them for a 240 k-point sweep.
// But RANSAC only needs the WINNER exactly.  In f32, |a x + b y + c z + d| is off by at most
// 4.5 * 2^-24 * (|x| + |y| + |z| + |d|) (coefficients rounded once, three fused multiply-adds), so with
// g = 8 * 2^-24 * (max_i (|x| + |y| + |z|) + |d|) the points below thr + g are an UPPER bound hi of a
// hypothesis' exact count and the points below thr - g a LOWER bound.  Every hypothesis gets its upper
// bound (k_ransac_upper: the whole batch, 4.25 instructions per test); the one with the largest upper bound
// gets its lower bound lo* (k_ransac_lo_seed, one thread per point).  The winner's exact count is at least
// lo*, so only hypotheses with hi >= lo* can win; those — a handful on a scene with a ground plane — are
// re-scored exactly.  Winner, tie-breaks, plane and mask are those of the full fp64 pass.
// also zeroes the upper-bound counters k_ransac_upper adds into
__global__ void __launch_bounds__(256) k_ransac_extent(const float4* __restrict__ pts, int64_t n,
                                                       unsigned* __restrict__ smax_bits, int32_t* __restrict__ hi,
                                                       int iters) {
    float m = 0.f;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < iters; i += stride) hi[i] = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 p = pts[i];
        const float v = fabsf(p.x) + fabsf(p.y) + fabsf(p.z);
        m = v > m ? v : m;           // a NaN coordinate never raises m: such points are outliers of every plane
        if (!(v < 3.0e38f)) m = 3.0e38f;   // an infinite / NaN extent disables the bounds (g becomes huge)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(smax_bits, __float_as_uint(m));   // non-negative floats order like their bits
}

// two IEEE fp32 fused multiply-adds per issue slot (sm_100 FFMA2); each half rounds like fmaf
__device__ __forceinline__ float2 rs_fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(r))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
          "l"(reinterpret_cast<unsigned long long&>(c)));
    return r;
}

// cnt += |t| < lim as one compare and one predicated add (the compiler's select form takes three instructions)
__device__ __forceinline__ void rs_count_below(int& cnt, float t, float lim) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "f"(fabsf(t)), "f"(lim));
}

// the f32 distance both bound kernels use (the error bound above is for exactly this nesting)
__device__ __forceinline__ float rs_dist32(float a, float b, float c, float d, float x, float y, float z) {
    return fabsf(fmaf(a, x, fmaf(b, y, fmaf(c, z, d))));
}
__device__ __forceinline__ double rs_guard(const unsigned* smax_bits, double dd) {
    return 8.0 * 5.9604644775390625e-08 * (static_cast<double>(__uint_as_float(*smax_bits)) + fabs(dd));
}

// hi[hyp] += this tile's count (integer atomics: order-independent).  A thread = a hypothesis; four points per
// trip: three 16-byte broadcast loads, six packed multiply-adds, four compares.  The tile is padded to a
// multiple of four with points at infinity (never inliers).
__global__ void __launch_bounds__(RS_THREADS) k_ransac_upper(const float4* __restrict__ pts, int64_t n, int flip_x,
                                                             const double* __restrict__ planes, int iters, double thr,
                                                             const unsigned* __restrict__ smax_bits,
                                                             int32_t* __restrict__ hi_out) {
    __shared__ __align__(16) float sx[RS_TILE], sy[RS_TILE], sz[RS_TILE];
    const int64_t p0 = static_cast<int64_t>(blockIdx.x) * RS_TILE;
    const int np = static_cast<int>(min(static_cast<int64_t>(RS_TILE), n - p0));
    const int np4 = (np + 3) & ~3;
    for (int i = threadIdx.x; i < np4; i += RS_THREADS) {
        float4 p = make_float4(INFINITY, INFINITY, INFINITY, 0.f);
        if (i < np) p = pts[p0 + i];
        sx[i] = flip_x ? -p.x : p.x, sy[i] = p.y, sz[i] = p.z;
    }
    __syncthreads();
    const int hyp = blockIdx.y * RS_THREADS + threadIdx.x;
    if (hyp >= iters) return;
    const double ad = planes[4 * hyp], bd = planes[4 * hyp + 1], cd = planes[4 * hyp + 2], dd = planes[4 * hyp + 3];
    int hi = 0;
    if (ad != 0.0 || bd != 0.0 || cd != 0.0 || dd != 0.0) {
        const float a = static_cast<float>(ad), b = static_cast<float>(bd), c = static_cast<float>(cd),
                    d = static_cast<float>(dd);
        const float t_hi = __double2float_ru(thr + rs_guard(smax_bits, dd));   // rounded outwards
        const float2 A = make_float2(a, a), B = make_float2(b, b), C = make_float2(c, c), D = make_float2(d, d);
        int hi2 = 0;   // two counters: the predicated adds form two chains instead of one
#pragma unroll 4
        for (int i = 0; i < np4; i += 4) {
            const float4 X = *reinterpret_cast<const float4*>(sx + i), Y = *reinterpret_cast<const float4*>(sy + i),
                         Z = *reinterpret_cast<const float4*>(sz + i);
            const float2 t01 = rs_fma2(A, make_float2(X.x, X.y),
                                       rs_fma2(B, make_float2(Y.x, Y.y), rs_fma2(C, make_float2(Z.x, Z.y), D)));
            const float2 t23 = rs_fma2(A, make_float2(X.z, X.w),
                                       rs_fma2(B, make_float2(Y.z, Y.w), rs_fma2(C, make_float2(Z.z, Z.w), D)));
            rs_count_below(hi, t01.x, t_hi);
            rs_count_below(hi2, t01.y, t_hi);
            rs_count_below(hi, t23.x, t_hi);
            rs_count_below(hi2, t23.y, t_hi);
        }
        hi += hi2;
    }
    if (hi) atomicAdd(hi_out + hyp, hi);
}

// the hypothesis with the largest upper bound (ties: smallest index)
__global__ void __launch_bounds__(256) k_ransac_seed(const int32_t* __restrict__ hi, int iters,
                                                     unsigned long long* __restrict__ seed_key) {
    const int hyp = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0ull;
    if (hyp < iters)
        key = (static_cast<unsigned long long>(hi[hyp]) << 32) | (0xffffffffu - static_cast<unsigned>(hyp));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(seed_key, key);
}

// lower bound of the seed hypothesis' exact inlier count: one thread per point
__global__ void __launch_bounds__(256) k_ransac_lo_seed(const float4* __restrict__ pts, int64_t n, int flip_x,
                                                        const double* __restrict__ planes, double thr,
                                                        const unsigned* __restrict__ smax_bits,
                                                        const unsigned long long* __restrict__ seed_key,
                                                        int32_t* __restrict__ lo_seed) {
    const unsigned long long key = *seed_key;
    if ((key >> 32) == 0ull) return;   // no hypothesis has an inlier: lo stays 0
    const int hyp = static_cast<int>(0xffffffffu - static_cast<unsigned>(key & 0xffffffffull));
    const double ad = planes[4 * hyp], bd = planes[4 * hyp + 1], cd = planes[4 * hyp + 2], dd = planes[4 * hyp + 3];
    const float a = static_cast<float>(ad), b = static_cast<float>(bd), c = static_cast<float>(cd),
                d = static_cast<float>(dd);
    const float t_lo = __double2float_rd(thr - rs_guard(smax_bits, dd));       // rounded outwards
    int lo = 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 p = pts[i];
        lo += rs_dist32(a, b, c, d, flip_x ? -p.x : p.x, p.y, p.z) < t_lo;
    }
    lo = __reduce_add_sync(0xffffffffu, lo);
    if ((threadIdx.x & 31) == 0 && lo) atomicAdd(lo_seed, lo);
}

// candidates: every hypothesis whose upper bound reaches the seed's lower bound (any order)
__global__ void __launch_bounds__(256) k_ransac_candidates(const int32_t* __restrict__ hi, const int32_t* __restrict__ lo_max,
                                                           int iters, int32_t* __restrict__ list,
                                                           int32_t* __restrict__ n_list, int32_t* __restrict__ cnt,
                                                           double* __restrict__ err) {
    const int hyp = blockIdx.x * blockDim.x + threadIdx.x;
    if (hyp >= iters) return;
    cnt[hyp] = 0;      // a hypothesis that cannot win takes no part in the arg-max
    err[hyp] = 0.0;
    if (hi[hyp] > 0 && hi[hyp] >= *lo_max) list[atomicAdd(n_list, 1)] = hyp;
}

// Sum of the per-tile partials of one hypothesis.  Order: 32 chunks of consecutive tiles, each summed in tile
// order, the chunk sums added in chunk order — the same bits from one thread (all hypotheses) and from a warp
// with a lane per chunk (the candidates).
// exact partials of the candidates (indexed by slot) -> cnt / err of their hypotheses; a warp per slot
__global__ void __launch_bounds__(256) k_ransac_reduce_list(const int32_t* __restrict__ part_cnt,
                                                            const double* __restrict__ part_err, int n_tiles, int iters,
                                                            const int32_t* __restrict__ list,
                                                            const int32_t* __restrict__ n_list, int32_t* __restrict__ cnt,
                                                            double* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int n_slots = min(*n_list, iters);
    const int per = (n_tiles + 31) / 32;
    const int t0 = lane * per, t1 = min(t0 + per, n_tiles);
    for (int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); slot < n_slots;
         slot += gridDim.x * (blockDim.x >> 5)) {
        int c = 0;
        double s = 0.0;
        for (int t = t0; t < t1; ++t) {
            c += part_cnt[static_cast<size_t>(t) * iters + slot];
            s += part_err[static_cast<size_t>(t) * iters + slot];
        }
        c = __reduce_add_sync(0xffffffffu, c);
        double e = 0.0;
        for (int l = 0; l < 32; ++l) e += __shfl_sync(0xffffffffu, s, l);
        if (lane == 0) {
            cnt[list[slot]] = c;
            err[list[slot]] = e;
        }
    }
}

__global__ void __launch_bounds__(256) k_ransac_reduce(const int32_t* __restrict__ part_cnt,
                                                       const double* __restrict__ part_err, int n_tiles, int iters,
                                                       int32_t* __restrict__ cnt, double* __restrict__ err) {
    const int hyp = blockIdx.x * blockDim.x + threadIdx.x;
    if (hyp >= iters) return;
    const int per = (n_tiles + 31) / 32;
    int c = 0;
    double e = 0.0;
    for (int t0 = 0; t0 < n_tiles; t0 += per) {
        const int t1 = min(t0 + per, n_tiles);
        double s = 0.0;
        for (int t = t0; t < t1; ++t) {
            c += part_cnt[static_cast<size_t>(t) * iters + hyp];
            s += part_err[static_cast<size_t>(t) * iters + hyp];
        }
        e += s;
    }
    cnt[hyp] = c;
    err[hyp] = e;
}

struct Cand {
    int cnt;
    double rmse;
    int idx;
};
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {
    // is a better than b?
    if (a.cnt != b.cnt) return a.cnt > b.cnt;
    if (a.rmse != b.rmse) return a.rmse < b.rmse;
    return a.idx < b.idx;
}

// one CTA; best: {index, count}; plane: winner
__global__ void __launch_bounds__(1024) k_ransac_best(const int32_t* __restrict__ cnt, const double* __restrict__ err,
                                                      const double* __restrict__ planes, int iters,
                                                      int32_t* __restrict__ best, double* __restrict__ plane) {
    Cand me{0, 1e300, 0x7fffffff};
    for (int i = threadIdx.x; i < iters; i += blockDim.x) {
        const int c = cnt[i];
        if (c <= 0) continue;
        Cand o{c, err[i] / sqrt(static_cast<double>(c)), i};
        if (better(o, me)) me = o;
    }
    __shared__ Cand s_c[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cand t;
        t.cnt = __shfl_down_sync(0xffffffffu, me.cnt, o);
        t.rmse = __shfl_down_sync(0xffffffffu, me.rmse, o);
        t.idx = __shfl_down_sync(0xffffffffu, me.idx, o);
        if (better(t, me)) me = t;
    }
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = me;
    __syncthreads();
    if (threadIdx.x == 0) {
        Cand b = s_c[0];
        for (int i = 1; i < static_cast<int>(blockDim.x >> 5); ++i)
            if (better(s_c[i], b)) b = s_c[i];
        const bool ok = b.cnt > 0;
        best[0] = ok ? b.idx : -1;
        best[1] = ok ? b.cnt : 0;
        for (int j = 0; j < 4; ++j) plane[j] = ok ? planes[4 * static_cast<size_t>(b.idx) + j] : 0.0;
    }
}

// inlier mask of the winner + one-pass moments (n, sum p, sum p p^T) for the refit.  No atomics: every CTA
// writes its ten partial moments (threads in a fixed grid-stride assignment, warp tree, warps in order) and
// k_ransac_refit adds the CTAs' partials in a fixed order — the refit is bit-reproducible.
constexpr int RS_MOM = 10;
template <int LAYOUT>
__global__ void __launch_bounds__(256) k_ransac_mask(const void* __restrict__ pts, int64_t n, int flip_x,
                                                     const double* __restrict__ plane, double thr,
                                                     uint8_t* __restrict__ mask, double* __restrict__ part_mom) {
    const double a = plane[0], b = plane[1], c = plane[2], d = plane[3];
    const bool have = a != 0.0 || b != 0.0 || c != 0.0 || d != 0.0;
    double m[RS_MOM] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double x, y, z;
        load_point<LAYOUT>(pts, i, flip_x, x, y, z);
        const bool in = have && plane_dist(a, b, c, d, x, y, z) < thr;
        mask[i] = in;
        if (in) {
            m[0] += 1.0, m[1] += x, m[2] += y, m[3] += z;
            m[4] += x * x, m[5] += x * y, m[6] += x * z, m[7] += y * y, m[8] += y * z, m[9] += z * z;
        }
    }
    __shared__ double s_m[8][RS_MOM];
#pragma unroll
    for (int j = 0; j < RS_MOM; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m[j] += __shfl_down_sync(0xffffffffu, m[j], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int j = 0; j < RS_MOM; ++j) s_m[threadIdx.x >> 5][j] = m[j];
    }
    __syncthreads();
    if (threadIdx.x < RS_MOM) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_m[w][threadIdx.x];
        part_mom[static_cast<size_t>(blockIdx.x) * RS_MOM + threadIdx.x] = t;
    }
}

// one warp: lane l adds the partials of the CTAs l, l + 32, .. in order, lane 0 adds the lanes' sums in lane order
__global__ void __launch_bounds__(32) k_ransac_refit(const double* __restrict__ part_mom, int n_parts,
                                                     const double* __restrict__ plane, double* __restrict__ refit) {
    const int lane = threadIdx.x;
    double mom[RS_MOM];
#pragma unroll
    for (int j = 0; j < RS_MOM; ++j) {
        double s = 0.0;
        for (int p = lane; p < n_parts; p += 32) s += part_mom[static_cast<size_t>(p) * RS_MOM + j];
        double t = 0.0;
        for (int l = 0; l < 32; ++l) t += __shfl_sync(0xffffffffu, s, l);
        mom[j] = t;
    }
    if (lane != 0) return;
    const double n = mom[0];
    if (n < 3.0) {
        for (int j = 0; j < 4; ++j) refit[j] = plane[j];
        return;
    }
    const double cx = mom[1] / n, cy = mom[2] / n, cz = mom[3] / n;
    double out[4];
    plane_from_moments(cx, cy, cz, mom[4] - n * cx * cx, mom[5] - n * cx * cy, mom[6] - n * cx * cz,
                       mom[7] - n * cy * cy, mom[8] - n * cy * cz, mom[9] - n * cz * cz, out);
    const bool zero = out[0] == 0.0 && out[1] == 0.0 && out[2] == 0.0;
    for (int j = 0; j < 4; ++j) refit[j] = zero ? plane[j] : out[j];
}

constexpr int RS_MASK_MAX_CTAS = 1024;   // k_ransac_mask grid cap (grid-stride beyond it)
struct RsWs {
    double* planes;
    int32_t* part_cnt;
    double* part_err;
    int32_t* cnt;
    double* err;
    double* mom;
    int32_t* hi;
    int32_t* list;
    int32_t* scalars;   // [0] seed's lower bound, [1] candidate count, [2] bits of max |x|+|y|+|z|, [4..5] seed key (u64)
};

size_t rs_carve(Bump& bump, RsWs& ws, int64_t n, int iters) {
    const size_t n_tiles = static_cast<size_t>(ceil_div64(n, RS_TILE));
    ws.planes = bump.take<double>(4 * static_cast<size_t>(iters));
    ws.part_cnt = bump.take<int32_t>(n_tiles * iters);
    ws.part_err = bump.take<double>(n_tiles * iters);
    ws.cnt = bump.take<int32_t>(iters);
    ws.err = bump.take<double>(iters);
    ws.mom = bump.take<double>(static_cast<size_t>(RS_MASK_MAX_CTAS) * RS_MOM);   // per-CTA partial moments
    ws.hi = bump.take<int32_t>(iters);
    ws.list = bump.take<int32_t>(iters);
    ws.scalars = bump.take<int32_t>(8);
    return bump.off;
}

template <int LAYOUT>
int rs_run(datmo_ctx* h, const void* pts, int64_t n, int flip_x, double thr, int ransac_n, int iters, uint64_t seed,
           double* plane, double* refit, uint8_t* inlier_mask, int32_t* best, double* hyp_planes, int32_t* hyp_count,
           double* hyp_err, size_t ws_offset) {
    RsWs ws;
    Bump bump(h->ws + ws_offset);
    rs_carve(bump, ws, n, iters);
    double* planes = hyp_planes ? hyp_planes : ws.planes;
    int32_t* cnt = hyp_count ? hyp_count : ws.cnt;
    double* err = hyp_err ? hyp_err : ws.err;
    const int n_tiles = static_cast<int>(ceil_div64(n, RS_TILE));
    {
        LaunchScope ls(h, DATMO_TAG_RANSAC);
        k_ransac_hyp<LAYOUT><<<ceil_div(iters, 128), 128, 0, h->stream>>>(pts, n, flip_x, ransac_n, iters, seed, planes);
    }
    DATMO_POST_LAUNCH(h);
    const dim3 g_score(n_tiles, ceil_div(iters, RS_THREADS));
    // the f32 bounds pass needs float points and is only worth it when nobody asked for every hypothesis' score
    const bool bounded = LAYOUT == DATMO_PTS_F32_XYZW && !hyp_count && !hyp_err && iters >= 2 * RS_THREADS &&
                         !getenv("DATMO_RANSAC_EXACT");
    if (bounded) {
        const float4* p4 = static_cast<const float4*>(pts);
        unsigned* smax = reinterpret_cast<unsigned*>(ws.scalars + 2);
        unsigned long long* seed_key = reinterpret_cast<unsigned long long*>(ws.scalars + 4);
        DATMO_CHECK_CUDA(h, cudaMemsetAsync(ws.scalars, 0, 8 * sizeof(int32_t), h->stream));
        {
            LaunchScope ls(h, DATMO_TAG_RANSAC);
            int64_t blocks = ceil_div64(n, 1024);
            k_ransac_extent<<<static_cast<int>(blocks > h->sm_count * 4 ? h->sm_count * 4 : blocks), 256, 0, h->stream>>>(p4, n, smax, ws.hi, iters);
        }
        DATMO_POST_LAUNCH(h);
        {
            LaunchScope ls(h, DATMO_TAG_RANSAC);
            k_ransac_upper<<<g_score, RS_THREADS, 0, h->stream>>>(p4, n, flip_x, planes, iters, thr, smax, ws.hi);
        }
        DATMO_POST_LAUNCH(h);
        {
            LaunchScope ls(h, DATMO_TAG_RANSAC);
            k_ransac_seed<<<ceil_div(iters, 256), 256, 0, h->stream>>>(ws.hi, iters, seed_key);
        }
        DATMO_POST_LAUNCH(h);
        {
            LaunchScope ls(h, DATMO_TAG_RANSAC);
            int64_t blocks = ceil_div64(n, 256);
            k_ransac_lo_seed<<<static_cast<int>(blocks > h->sm_count * 8 ? h->sm_count * 8 : blocks), 256, 0, h->stream>>>(
                p4, n, flip_x, planes, thr, smax, seed_key, ws.scalars);
        }
        DATMO_POST_LAUNCH(h);
        {
            LaunchScope ls(h, DATMO_TAG_RANSAC);
            k_ransac_candidates<<<ceil_div(iters, 256), 256, 0, h->stream>>>(ws.hi, ws.scalars, iters, ws.list, ws.scalars + 1, cnt, err);
        }
        DATMO_POST_LAUNCH(h);
        {
            LaunchScope ls(h, DATMO_TAG_RANSAC);
            k_ransac_score<LAYOUT><<<dim3(n_tiles, 1), RS_THREADS, 0, h->stream>>>(pts, n, flip_x, planes, iters, thr,
                                                                                   ws.part_cnt, ws.part_err, ws.list,
                                                                                   ws.scalars + 1);
        }
        DATMO_POST_LAUNCH(h);
        {
            LaunchScope ls(h, DATMO_TAG_RANSAC);
            k_ransac_reduce_list<<<std::min(ceil_div(iters, 8), 4 * h->sm_count), 256, 0, h->stream>>>(
                ws.part_cnt, ws.part_err, n_tiles, iters, ws.list, ws.scalars + 1, cnt, err);
        }
        DATMO_POST_LAUNCH(h);
    } else {
        {
            LaunchScope ls(h, DATMO_TAG_RANSAC);
            k_ransac_score<LAYOUT><<<g_score, RS_THREADS, 0, h->stream>>>(pts, n, flip_x, planes, iters, thr, ws.part_cnt,
                                                                          ws.part_err, nullptr, nullptr);
        }
        DATMO_POST_LAUNCH(h);
        {
            LaunchScope ls(h, DATMO_TAG_RANSAC);
            k_ransac_reduce<<<ceil_div(iters, 256), 256, 0, h->stream>>>(ws.part_cnt, ws.part_err, n_tiles, iters, cnt, err);
        }
        DATMO_POST_LAUNCH(h);
    }
    {
        LaunchScope ls(h, DATMO_TAG_RANSAC);
        k_ransac_best<<<1, 1024, 0, h->stream>>>(cnt, err, planes, iters, best, plane);
    }
    DATMO_POST_LAUNCH(h);
    int64_t mask_blocks = ceil_div64(n, 256);
    const int mask_grid = static_cast<int>(std::min<int64_t>(mask_blocks, std::min(h->sm_count * 4, RS_MASK_MAX_CTAS)));
    {
        LaunchScope ls(h, DATMO_TAG_RANSAC);
        k_ransac_mask<LAYOUT><<<mask_grid, 256, 0, h->stream>>>(pts, n, flip_x, plane, thr, inlier_mask, ws.mom);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, DATMO_TAG_RANSAC);
        k_ransac_refit<<<1, 32, 0, h->stream>>>(ws.mom, mask_grid, plane, refit);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

}  // namespace

size_t datmo_ransac_ws_bytes(int64_t n, int iters) {
    Bump dry(nullptr);
    RsWs ws;
    return rs_carve(dry, ws, n, iters);
}

// shared with bev.cu's fused preprocessing; scratch lives at h->ws + ws_offset (already reserved)
int datmo_ransac_run(datmo_ctx* h, const void* pts, int layout, int64_t n, int flip_x, double thr, int ransac_n,
                     int iters, uint64_t seed, double* plane, double* refit, uint8_t* inlier_mask, int32_t* best,
                     double* hyp_planes, int32_t* hyp_count, double* hyp_err, size_t ws_offset) {
    DATMO_REQUIRE(h, pts && plane && refit && inlier_mask && best, "null pointer");
    DATMO_REQUIRE(h, layout == DATMO_PTS_F64_XYZ || layout == DATMO_PTS_F32_XYZW, "unknown point layout");
    DATMO_REQUIRE(h, ransac_n >= 3 && ransac_n <= RS_MAX_N, "ransac_n must be in [3, 8]");
    DATMO_REQUIRE(h, n >= ransac_n, "There must be at least 'ransac_n' points.");
    DATMO_REQUIRE(h, iters >= 1 && iters <= (1 << 24) && thr > 0, "bad num_iterations / distance_threshold");
    if (layout == DATMO_PTS_F64_XYZ)
        return rs_run<DATMO_PTS_F64_XYZ>(h, pts, n, flip_x, thr, ransac_n, iters, seed, plane, refit, inlier_mask, best,
                                         hyp_planes, hyp_count, hyp_err, ws_offset);
    return rs_run<DATMO_PTS_F32_XYZW>(h, pts, n, flip_x, thr, ransac_n, iters, seed, plane, refit, inlier_mask, best,
                                      hyp_planes, hyp_count, hyp_err, ws_offset);
}

extern "C" int datmo_ransac_ground_dev(datmo_handle_t h, const void* pts, int layout, int64_t n, int flip_x,
                                       double distance_threshold, int ransac_n, int num_iterations, uint64_t seed,
                                       double* plane, double* refit, uint8_t* inlier_mask, int32_t* best,
                                       double* hyp_planes, int32_t* hyp_count, double* hyp_err) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, n >= 1 && num_iterations >= 1, "empty cloud / no iterations");
    DATMO_TRY(datmo_ws_reserve(h, datmo_ransac_ws_bytes(n, num_iterations)));
    return datmo_ransac_run(h, pts, layout, n, flip_x, distance_threshold, ransac_n, num_iterations, seed, plane, refit,
                            inlier_mask, best, hyp_planes, hyp_count, hyp_err, 0);
}
