// LiDAR cloud -> BEV height grid, and the fused per-frame preprocessing.  sm_100a.
//
// Replaces compute_bev_grid (Optical_flow/main.py:98-126) and, fused in front of it,
// the flip / ground drop / ROI crop / density expansion of preprocess_pcd
// (main.py:65-92, with filter_points_in_roi main.py:30-36 and increase_point_density
// main.py:38-57).
//
// Bit-exactness with the reference's uint8 grid needs fp64 index math with truncation
// toward zero ((x - lo) / w, so x in (lo - w, lo) lands in cell 0), fp64 per-cell
// statistics (two passes: mean, then squared deviations — np.mean / np.std), every
// operation of the value formula rounded separately (no FMA), and numpy's
// float64 -> uint8 cast (truncate through int32, out of range -> INT_MIN, low byte).
// Accumulation is an atomic scatter: warp-aggregated (the x-expansion copies of one
// return are adjacent and almost always share a cell) into fp64 global accumulators
// that stay L2-resident (24 B / cell).
#include <math.h>

#include "common.cuh"

namespace {

struct BevGeom {
    double x_lo, y_lo, res_x, res_y;
    int nx, ny;
};

__device__ __forceinline__ int cell_of(double x, double y, const BevGeom& g) {
    // main.py:106-108: int((x - lo) / w), keep if 0 <= idx < nbins
    const double qx = trunc(__ddiv_rn(__dsub_rn(x, g.x_lo), g.res_x));
    const double qy = trunc(__ddiv_rn(__dsub_rn(y, g.y_lo), g.res_y));
    if (!(qx >= 0.0 && qx < static_cast<double>(g.nx) && qy >= 0.0 && qy < static_cast<double>(g.ny))) return -1;
    return static_cast<int>(qx) * g.ny + static_cast<int>(qy);
}

__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// two N(0,1) draws from one counter (Box-Muller)
__device__ __forceinline__ void normal2(uint64_t seed, uint64_t ctr, double& n0, double& n1) {
    uint64_t a = mix64(seed, 2 * ctr), b = mix64(seed, 2 * ctr + 1);
    double u0 = (static_cast<double>(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0, 1]
    double u1 = static_cast<double>(b >> 11) * (1.0 / 9007199254740992.0);
    double rad = sqrt(-2.0 * log(u0));
    double s, c;
    sincospi(2.0 * u1, &s, &c);
    n0 = rad * c;
    n1 = rad * s;
}

// Per-cell accumulator that does not depend on the order in which warps and CTAs arrive: a 128-bit
// fixed-point sum (hi = integer part as int64, lo = fraction in units of 2^-64; see datmo_fixed_add in
// common.cuh) plus an fp64 accumulator for non-finite input.  With fp64 atomicAdd the last bits of a
// cell's mean changed from run to run, and with them — rarely — a uint8 of the grid.
struct FixAcc {
    unsigned long long* hi;
    unsigned long long* lo;
    double* ovf;
};
__device__ __forceinline__ double fix_value(const FixAcc& a, int i) {
    return datmo_fixed_value(a.hi[i], a.lo[i], a.ovf[i]);
}

// Adds (count, value) for `cell` with one set of atomics per distinct cell in the warp.
// Lanes with cell < 0 contribute nothing.  All 32 lanes must call.
__device__ __forceinline__ void warp_scatter_add(int cell, double v, uint32_t* cnt, const FixAcc& acc) {
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    // sum v over the peer group in lane order (deterministic within the warp)
    double sum = 0.0;
    unsigned rem = peers;
    while (__any_sync(0xffffffffu, rem != 0)) {
        int src = rem ? __ffs(rem) - 1 : 0;
        double t = __shfl_sync(0xffffffffu, v, src);
        if (rem) {
            sum += t;
            rem &= rem - 1;
        }
    }
    if (lane == leader && cell >= 0) {
        if (cnt) atomicAdd(cnt + cell, static_cast<uint32_t>(__popc(peers)));
        datmo_fixed_add(acc.hi + cell, acc.lo + cell, reinterpret_cast<unsigned long long*>(acc.ovf + cell), sum);
    }
}

template <int LAYOUT>
__device__ __forceinline__ void load_point(const void* pts, int64_t i, double& x, double& y, double& z) {
    if (LAYOUT == DATMO_PTS_F64_XYZ) {
        const double* p = static_cast<const double*>(pts) + 3 * i;
        x = p[0], y = p[1], z = p[2];
    } else {
        const float4 p = static_cast<const float4*>(pts)[i];
        x = p.x, y = p.y, z = p.z;
    }
}

// pass 1 (PASS == 0): count and sum of z per cell; pass 2 (PASS == 1): sum of squared deviations
template <int LAYOUT, int PASS>
__global__ void __launch_bounds__(256) k_bev_accum(const void* __restrict__ pts, int64_t n, BevGeom g,
                                                   uint32_t* __restrict__ cnt, FixAcc sum, FixAcc ssd) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t n_round = (n + 31) & ~int64_t(31);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        int cell = -1;
        double v = 0.0;
        if (i < n) {
            double x, y, z;
            load_point<LAYOUT>(pts, i, x, y, z);
            cell = cell_of(x, y, g);
            if (cell >= 0) {
                if (PASS == 0) {
                    v = z;
                } else {
                    const double mean = __ddiv_rn(fix_value(sum, cell), static_cast<double>(cnt[cell]));
                    const double d = __dsub_rn(z, mean);
                    v = __dmul_rn(d, d);
                }
            }
        }
        warp_scatter_add(cell, v, PASS == 0 ? cnt : nullptr, PASS == 0 ? sum : ssd);
    }
}

// Fused preprocessing scatter: flip, ground drop, ROI crop (on the un-noised point),
// `expansion` noisy copies, cell scatter.  One thread per (point, copy) so the copies
// of a point sit in adjacent lanes and aggregate in warp_scatter_add.
template <int PASS>
__global__ void __launch_bounds__(256) k_pre_accum(const float4* __restrict__ pts, int64_t n, int flip_x,
                                                   const uint8_t* __restrict__ ground, double rx0, double rx1,
                                                   double ry0, double ry1, double rz0, double rz1, int expansion,
                                                   double noise_std, const double* __restrict__ noise, uint64_t seed,
                                                   BevGeom g, uint32_t* __restrict__ cnt, FixAcc sum, FixAcc ssd,
                                                   unsigned long long* __restrict__ n_roi) {
    const int64_t total = n * expansion;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t_round = (total + 31) & ~int64_t(31);
    unsigned long long roi_local = 0;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < t_round; t += stride) {
        int cell = -1;
        double v = 0.0;
        if (t < total) {
            const int64_t i = t / expansion;
            const int e = static_cast<int>(t - i * expansion);
            const float4 p = pts[i];
            double x = flip_x ? -static_cast<double>(p.x) : static_cast<double>(p.x);
            double y = p.y, z = p.z;
            const bool keep = !(ground && ground[i]) && x >= rx0 && x <= rx1 && y >= ry0 && y <= ry1 && z >= rz0 &&
                              z <= rz1;
            if (keep) {
                if (PASS == 0 && e == 0) ++roi_local;
                double nx_, ny_, nz_;
                if (noise) {
                    const double* q = noise + (i * expansion + e) * 3;
                    nx_ = q[0], ny_ = q[1], nz_ = q[2];
                } else {
                    double a, b, c, d;
                    normal2(seed, 2 * static_cast<uint64_t>(t), a, b);
                    normal2(seed, 2 * static_cast<uint64_t>(t) + 1, c, d);
                    nx_ = a * noise_std, ny_ = b * noise_std, nz_ = c * noise_std;
                }
                x = __dadd_rn(x, nx_), y = __dadd_rn(y, ny_), z = __dadd_rn(z, nz_);
                cell = cell_of(x, y, g);
                if (cell >= 0) {
                    if (PASS == 0) {
                        v = z;
                    } else {
                        const double mean = __ddiv_rn(fix_value(sum, cell), static_cast<double>(cnt[cell]));
                        const double d = __dsub_rn(z, mean);
                        v = __dmul_rn(d, d);
                    }
                }
            }
        }
        warp_scatter_add(cell, v, PASS == 0 ? cnt : nullptr, PASS == 0 ? sum : ssd);
    }
    if (PASS == 0 && n_roi) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) roi_local += __shfl_down_sync(0xffffffffu, roi_local, o);
        if ((threadIdx.x & 31) == 0 && roi_local) atomicAdd(n_roi, roi_local);
    }
}

// order-preserving map double <-> uint64 so atomicMax works on signed values
__device__ __forceinline__ unsigned long long ord_of(double v) {
    unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_to_double(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double(static_cast<long long>(u));
}

// per-cell value (a*mean + b*std)/h_max (main.py:116-118) -> val; grid max
__global__ void __launch_bounds__(256) k_bev_value(const uint32_t* __restrict__ cnt, FixAcc sum, FixAcc ssd,
                                                   double* __restrict__ val, int ncell, double a, double b,
                                                   double h_max, unsigned long long* __restrict__ gmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (i < ncell) {
        const uint32_t c = cnt[i];
        if (c > 0) {
            const double dn = static_cast<double>(c);
            const double mean = __ddiv_rn(fix_value(sum, i), dn);
            const double sd = sqrt(__ddiv_rn(fix_value(ssd, i), dn));
            v = __ddiv_rn(__dadd_rn(__dmul_rn(a, mean), __dmul_rn(b, sd)), h_max);
        }
        val[i] = v;
    }
    unsigned long long key = i < ncell ? ord_of(v) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_down_sync(0xffffffffu, key, o);
        key = t > key ? t : key;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(gmax, key);
}

// main.py:122-123: v / max * 255 -> uint8 with numpy's cast
__global__ void __launch_bounds__(256) k_bev_norm(const double* __restrict__ vals, int ncell,
                                                  const unsigned long long* __restrict__ gmax,
                                                  uint8_t* __restrict__ bev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    const double mx = ord_to_double(*gmax);
    const double s = __dmul_rn(__ddiv_rn(vals[i], mx), 255.0);
    uint8_t o = 0;
    if (s > -2147483649.0 && s < 2147483648.0) o = static_cast<uint8_t>(__double2int_rz(s) & 0xFF);
    bev[i] = o;
}

// ---- filter_points_in_roi (main.py:30-36) as a stable device compaction ----------------------
template <int LAYOUT>
__global__ void __launch_bounds__(256) k_roi_flags(const void* __restrict__ pts, int64_t n, double x0, double x1,
                                                   double y0, double y1, double z0, double z1,
                                                   uint8_t* __restrict__ flags) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z;
    load_point<LAYOUT>(pts, i, x, y, z);
    flags[i] = x >= x0 && x <= x1 && y >= y0 && y <= y1 && z >= z0 && z <= z1;   // closed intervals
}

template <int LAYOUT>
__global__ void __launch_bounds__(256) k_roi_scatter(const void* __restrict__ pts, int64_t n,
                                                     const uint8_t* __restrict__ flags,
                                                     const int32_t* __restrict__ rank, void* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const int64_t o = rank[i];
    if (LAYOUT == DATMO_PTS_F64_XYZ) {
        const double* p = static_cast<const double*>(pts) + 3 * i;
        double* q = static_cast<double*>(out) + 3 * o;
        q[0] = p[0], q[1] = p[1], q[2] = p[2];
    } else {
        static_cast<float4*>(out)[o] = static_cast<const float4*>(pts)[i];
    }
}

// ---- increase_point_density (main.py:38-57): consecutive copies + noise ---------------------------
__global__ void __launch_bounds__(256) k_expand(const double* __restrict__ pts, int64_t n, int expansion,
                                                double noise_std, const double* __restrict__ noise, uint64_t seed,
                                                double* __restrict__ out) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n * expansion) return;
    const int64_t i = t / expansion;
    double nx_, ny_, nz_;
    if (noise) {
        nx_ = noise[3 * t], ny_ = noise[3 * t + 1], nz_ = noise[3 * t + 2];
    } else {
        double a, b, c, d;
        normal2(seed, 2 * static_cast<uint64_t>(t), a, b);
        normal2(seed, 2 * static_cast<uint64_t>(t) + 1, c, d);
        nx_ = a * noise_std, ny_ = b * noise_std, nz_ = c * noise_std;
    }
    out[3 * t] = __dadd_rn(pts[3 * i], nx_);
    out[3 * t + 1] = __dadd_rn(pts[3 * i + 1], ny_);
    out[3 * t + 2] = __dadd_rn(pts[3 * i + 2], nz_);
}

struct BevWs {
    uint32_t* cnt;
    FixAcc sum, ssd;
    double* val;
    unsigned long long* gmax;
    unsigned long long* n_roi;
};

void bev_carve(Bump& bump, int ncell, BevWs& ws) {
    ws.cnt = bump.take<uint32_t>(ncell);
    for (FixAcc* a : {&ws.sum, &ws.ssd}) {
        a->hi = bump.take<unsigned long long>(ncell);
        a->lo = bump.take<unsigned long long>(ncell);
        a->ovf = bump.take<double>(ncell);
    }
    ws.val = bump.take<double>(ncell);
    ws.gmax = bump.take<unsigned long long>(2);
    ws.n_roi = ws.gmax + 1;
}

int bev_ws(datmo_ctx* h, int ncell, BevWs& ws) {
    for (int pass = 0; pass < 2; ++pass) {
        Bump bump(pass ? h->ws : nullptr);
        bev_carve(bump, ncell, ws);
        if (!pass) DATMO_TRY(datmo_ws_reserve(h, bump.off));
        if (pass) DATMO_CHECK_CUDA(h, cudaMemsetAsync(h->ws, 0, bump.off, h->stream));
    }
    return DATMO_OK;
}

int bev_finish(datmo_ctx* h, const BevWs& ws, int ncell, double a, double b, double h_max, uint8_t* bev) {
    {
        LaunchScope ls(h, DATMO_TAG_BEV);
        k_bev_value<<<ceil_div(ncell, 256), 256, 0, h->stream>>>(ws.cnt, ws.sum, ws.ssd, ws.val, ncell, a, b, h_max, ws.gmax);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, DATMO_TAG_BEV);
        k_bev_norm<<<ceil_div(ncell, 256), 256, 0, h->stream>>>(ws.val, ncell, ws.gmax, bev);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

int scatter_grid(datmo_ctx* h, int64_t items) {
    int64_t blocks = ceil_div64(items, 256);
    int64_t cap = static_cast<int64_t>(h->sm_count) * 16;  // a few waves of a grid-stride loop
    return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

template <int LAYOUT>
int bev_run(datmo_ctx* h, const void* pts, int64_t n, const BevGeom& g, double a, double b, double h_max,
            uint8_t* bev) {
    const int ncell = g.nx * g.ny;
    BevWs ws;
    DATMO_TRY(bev_ws(h, ncell, ws));
    if (n > 0) {
        int grid = scatter_grid(h, n);
        {
            LaunchScope ls(h, DATMO_TAG_BEV);
            k_bev_accum<LAYOUT, 0><<<grid, 256, 0, h->stream>>>(pts, n, g, ws.cnt, ws.sum, ws.ssd);
        }
        DATMO_POST_LAUNCH(h);
        {
            LaunchScope ls(h, DATMO_TAG_BEV);
            k_bev_accum<LAYOUT, 1><<<grid, 256, 0, h->stream>>>(pts, n, g, ws.cnt, ws.sum, ws.ssd);
        }
        DATMO_POST_LAUNCH(h);
    }
    return bev_finish(h, ws, ncell, a, b, h_max, bev);
}

int check_geom(datmo_ctx* h, double res_x, double res_y, int nx, int ny) {
    DATMO_REQUIRE(h, res_x > 0 && res_y > 0, "grid resolution must be positive");
    DATMO_REQUIRE(h, nx >= 1 && ny >= 1 && static_cast<int64_t>(nx) * ny < (int64_t(1) << 30), "bad grid size");
    return DATMO_OK;
}

}  // namespace

// ransac.cu
int datmo_ransac_run(datmo_ctx* h, const void* pts, int layout, int64_t n, int flip_x, double thr, int ransac_n,
                     int iters, uint64_t seed, double* plane, double* refit, uint8_t* inlier_mask, int32_t* best,
                     double* hyp_planes, int32_t* hyp_count, double* hyp_err, size_t ws_offset);
size_t datmo_ransac_ws_bytes(int64_t n, int iters);

extern "C" {

int datmo_bev_bins(double lo, double hi, double step) {
    // len(np.arange(lo, hi, step)) = ceil((hi - lo) / step), numpy's _arange_safe_ceil_to_intp
    if (!(step > 0) || !(hi > lo)) return 0;
    double len = ceil((hi - lo) / step);
    if (!(len < 2147483647.0)) return -1;
    return static_cast<int>(len);
}

int datmo_bev_rasterize_dev(datmo_handle_t h, const void* pts, int layout, int64_t n, double res_x, double res_y,
                            double x_lo, double y_lo, int nx, int ny, double a, double b, double h_max, uint8_t* bev) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, (pts || n == 0) && bev && n >= 0, "null pointer");
    DATMO_REQUIRE(h, layout == DATMO_PTS_F64_XYZ || layout == DATMO_PTS_F32_XYZW, "unknown point layout");
    DATMO_TRY(check_geom(h, res_x, res_y, nx, ny));
    BevGeom g{x_lo, y_lo, res_x, res_y, nx, ny};
    if (layout == DATMO_PTS_F64_XYZ) return bev_run<DATMO_PTS_F64_XYZ>(h, pts, n, g, a, b, h_max, bev);
    return bev_run<DATMO_PTS_F32_XYZW>(h, pts, n, g, a, b, h_max, bev);
}

int datmo_bev_rasterize_host(datmo_handle_t h, const void* pts, int layout, int64_t n, double res_x, double res_y,
                             double x_lo, double y_lo, int nx, int ny, double a, double b, double h_max, uint8_t* bev) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, (pts || n == 0) && bev && n >= 0, "null pointer");
    DATMO_REQUIRE(h, layout == DATMO_PTS_F64_XYZ || layout == DATMO_PTS_F32_XYZW, "unknown point layout");
    DATMO_TRY(check_geom(h, res_x, res_y, nx, ny));
    const size_t pbytes = static_cast<size_t>(n) * (layout == DATMO_PTS_F64_XYZ ? 24 : 16);
    const size_t ncell = static_cast<size_t>(nx) * ny;
    DATMO_TRY(datmo_io_reserve(h, pbytes + ncell + 512));
    char* d = h->io;
    uint8_t* d_bev = reinterpret_cast<uint8_t*>(d + ((pbytes + 255) & ~size_t(255)));
    int st = DATMO_OK;
    cudaError_t e = cudaSuccess;
    if (pbytes && (e = cudaMemcpyAsync(d, pts, pbytes, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        st = DATMO_E_CUDA;
    }
    if (st == DATMO_OK) st = datmo_bev_rasterize_dev(h, d, layout, n, res_x, res_y, x_lo, y_lo, nx, ny, a, b, h_max, d_bev);
    if (st == DATMO_OK && (e = cudaMemcpyAsync(bev, d_bev, ncell, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        st = DATMO_E_CUDA;
    }
    e = cudaStreamSynchronize(h->stream);
    if (st == DATMO_OK && e != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        st = DATMO_E_CUDA;
    }
    return st;
}

int datmo_preprocess_dev(datmo_handle_t h, const float* pts, int64_t n, int flip_x, double distance_threshold,
                         int ransac_n, int num_iterations, uint64_t seed, const uint8_t* ground_mask,
                         const double roi[6], int expansion, double noise_std, const double* noise, double res_x,
                         double res_y, double x_lo, double y_lo, int nx, int ny, double h_max, uint8_t* bev,
                         int64_t* n_roi) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, pts && bev && roi && n >= 1, "null pointer / empty cloud");
    DATMO_REQUIRE(h, expansion >= 1 && expansion <= 1024, "expansion out of range");
    DATMO_TRY(check_geom(h, res_x, res_y, nx, ny));
    const int ncell = nx * ny;
    BevGeom g{x_lo, y_lo, res_x, res_y, nx, ny};
    // workspace: [bev accumulators][ground mask][ransac scratch]
    size_t bev_bytes;
    {
        Bump dry(nullptr);
        BevWs sizing;
        bev_carve(dry, ncell, sizing);
        bev_bytes = dry.off;
    }
    const size_t mask_bytes = (static_cast<size_t>(n) + 255) & ~size_t(255);
    const bool run_ransac = ground_mask == nullptr && num_iterations > 0;
    const size_t ransac_bytes = run_ransac ? datmo_ransac_ws_bytes(n, num_iterations) + 1024 : 0;
    DATMO_TRY(datmo_ws_reserve(h, bev_bytes + mask_bytes + ransac_bytes));
    BevWs ws;
    DATMO_TRY(bev_ws(h, ncell, ws));
    uint8_t* d_mask = reinterpret_cast<uint8_t*>(h->ws + bev_bytes);
    const uint8_t* ground = ground_mask;
    if (run_ransac) {
        double* d_plane = reinterpret_cast<double*>(h->ws + bev_bytes + mask_bytes);
        double* d_refit = d_plane + 4;
        int32_t* d_best = reinterpret_cast<int32_t*>(d_refit + 4);
        DATMO_TRY(datmo_ransac_run(h, pts, DATMO_PTS_F32_XYZW, n, flip_x, distance_threshold, ransac_n,
                                   num_iterations, seed, d_plane, d_refit, d_mask, d_best, nullptr, nullptr, nullptr,
                                   bev_bytes + mask_bytes + 1024));
        ground = d_mask;
    }
    const int grid = scatter_grid(h, n * expansion);
    const float4* p4 = reinterpret_cast<const float4*>(pts);
    {
        LaunchScope ls(h, DATMO_TAG_BEV);
        k_pre_accum<0><<<grid, 256, 0, h->stream>>>(p4, n, flip_x, ground, roi[0], roi[1], roi[2], roi[3], roi[4],
                                                    roi[5], expansion, noise_std, noise, seed ^ 0xD1B54A32D192ED03ull,
                                                    g, ws.cnt, ws.sum, ws.ssd, ws.n_roi);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, DATMO_TAG_BEV);
        k_pre_accum<1><<<grid, 256, 0, h->stream>>>(p4, n, flip_x, ground, roi[0], roi[1], roi[2], roi[3], roi[4],
                                                    roi[5], expansion, noise_std, noise, seed ^ 0xD1B54A32D192ED03ull,
                                                    g, ws.cnt, ws.sum, ws.ssd, nullptr);
    }
    DATMO_POST_LAUNCH(h);
    DATMO_TRY(bev_finish(h, ws, ncell, 0.5, 0.5, h_max, bev));
    unsigned long long roi_count = 0;
    DATMO_CHECK_CUDA(h, cudaMemcpyAsync(&roi_count, ws.n_roi, sizeof(roi_count), cudaMemcpyDeviceToHost, h->stream));
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (n_roi) *n_roi = static_cast<int64_t>(roi_count);
    return roi_count == 0 ? DATMO_E_EMPTY : DATMO_OK;
}

int datmo_roi_filter_dev(datmo_handle_t h, const void* pts, int layout, int64_t n, const double roi[6], void* out,
                         int64_t* n_out) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, (pts || n == 0) && roi && (out || n == 0) && n_out && n >= 0, "null pointer");
    DATMO_REQUIRE(h, layout == DATMO_PTS_F64_XYZ || layout == DATMO_PTS_F32_XYZW, "unknown point layout");
    DATMO_REQUIRE(h, n < (int64_t(1) << 31), "too many points");
    *n_out = 0;
    if (n == 0) return DATMO_OK;
    const int nblk = static_cast<int>(ceil_div64(n, 4096));
    uint8_t* flags;
    int32_t *rank, *bsum, *total;
    for (int pass = 0; pass < 2; ++pass) {
        Bump bump(pass ? h->ws : nullptr);
        flags = bump.take<uint8_t>(n);
        rank = bump.take<int32_t>(n);
        bsum = bump.take<int32_t>(nblk);
        total = bump.take<int32_t>(1);
        if (!pass) DATMO_TRY(datmo_ws_reserve(h, bump.off));
    }
    const unsigned g = static_cast<unsigned>(ceil_div64(n, 256));
    {
        LaunchScope ls(h, DATMO_TAG_BEV);
        if (layout == DATMO_PTS_F64_XYZ)
            k_roi_flags<DATMO_PTS_F64_XYZ><<<g, 256, 0, h->stream>>>(pts, n, roi[0], roi[1], roi[2], roi[3], roi[4],
                                                                    roi[5], flags);
        else
            k_roi_flags<DATMO_PTS_F32_XYZW><<<g, 256, 0, h->stream>>>(pts, n, roi[0], roi[1], roi[2], roi[3], roi[4],
                                                                     roi[5], flags);
    }
    DATMO_POST_LAUNCH(h);
    DATMO_TRY(datmo_flag_scan(h, flags, n, 1, bsum, total, rank, DATMO_TAG_BEV, 1));
    {
        LaunchScope ls(h, DATMO_TAG_BEV);
        if (layout == DATMO_PTS_F64_XYZ)
            k_roi_scatter<DATMO_PTS_F64_XYZ><<<g, 256, 0, h->stream>>>(pts, n, flags, rank, out);
        else
            k_roi_scatter<DATMO_PTS_F32_XYZW><<<g, 256, 0, h->stream>>>(pts, n, flags, rank, out);
    }
    DATMO_POST_LAUNCH(h);
    int32_t cnt = 0;
    DATMO_CHECK_CUDA(h, cudaMemcpyAsync(&cnt, total, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    *n_out = cnt;
    return DATMO_OK;
}

int datmo_expand_points_dev(datmo_handle_t h, const double* pts, int64_t n, int expansion, double noise_std,
                            const double* noise, uint64_t seed, double* out) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, (pts || n == 0) && out && n >= 0 && expansion >= 1, "bad arguments");
    if (n == 0) return DATMO_OK;
    {
        LaunchScope ls(h, DATMO_TAG_BEV);
        k_expand<<<static_cast<unsigned>(ceil_div64(n * expansion, 256)), 256, 0, h->stream>>>(pts, n, expansion,
                                                                                              noise_std, noise, seed, out);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

}  // extern "C"
