// Velocity grid, continuity mask and moving-cell filter, sm_100a.
//
// Replaces, in one pass over the flow field:
//   compute_velocity_vectors tail   Optical_flow/main.py:143-164  velocity = flow * pixel size, curl
//   continuity_mask                 Optical_flow/main.py:224-228  |div| <= a and |curl| <= a (np.gradient)
//   inline moving-cell filter       Optical_flow/main.py:596-609  v * mask, magnitude, mag > 0.1
// np.gradient: central difference / 2 in the interior, one-sided first order at the
// edges, f32 in -> f32 out; the comparison against alpha happens in f32 (numpy's weak
// python-scalar promotion); v * mask promotes to f64 and the magnitude test runs in f64.
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ float grad1(float lo, float c, float hi, int i, int n) {
    // lo / hi are the samples at i-1 / i+1 (unused at the edges)
    if (i == 0) return hi - c;
    if (i == n - 1) return c - lo;
    return (hi - lo) / 2.0f;
}

// A thread owns one column and VM_ROWS consecutive rows: the column's flow is loaded once for all of
// them (VM_ROWS + 2 loads instead of 3 per row), and the CTA count drops by VM_ROWS — with one
// pixel per thread the kernel was bound by CTA turnover, not by bandwidth.
constexpr int VM_ROWS = 4;

__global__ void __launch_bounds__(256) k_velmask(const float2* __restrict__ flow, int H, int W, float px, float py,
                                                 float alpha, double s_crit, float s_lo, float s_hi,
                                                 float* __restrict__ vx_o,
                                                 float* __restrict__ vy_o, float* __restrict__ ang_o,
                                                 uint8_t* __restrict__ mask_o, float* __restrict__ vxf_o,
                                                 float* __restrict__ vyf_o, uint8_t* __restrict__ valid_o,
                                                 int32_t* __restrict__ n_valid) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * VM_ROWS, b = blockIdx.z;
    int n_ok = 0;
    if (x < W) {
        const size_t base = static_cast<size_t>(b) * H * W;
        const float2* f = flow + base;
        float2 col[VM_ROWS + 2], lft[VM_ROWS], rgt[VM_ROWS];
#pragma unroll
        for (int k = 0; k < VM_ROWS + 2; ++k) col[k] = f[static_cast<size_t>(min(max(y0 - 1 + k, 0), H - 1)) * W + x];
#pragma unroll
        for (int k = 0; k < VM_ROWS; ++k) {
            const size_t o = static_cast<size_t>(min(y0 + k, H - 1)) * W + x;
            lft[k] = f[o - (x > 0)];
            rgt[k] = f[o + (x < W - 1)];
        }
#pragma unroll
        for (int k = 0; k < VM_ROWS; ++k) {
            const int y = y0 + k;
            if (y >= H) break;
            const size_t o = static_cast<size_t>(y) * W + x;
            const float2 c = col[k + 1], u = col[k], d = col[k + 2], l = lft[k], r = rgt[k];
            const float vx = __fmul_rn(c.x, px), vy = __fmul_rn(c.y, py);
            const float dvx_dx = grad1(__fmul_rn(l.x, px), vx, __fmul_rn(r.x, px), x, W);
            const float dvy_dx = grad1(__fmul_rn(l.y, py), vy, __fmul_rn(r.y, py), x, W);
            const float dvx_dy = grad1(__fmul_rn(u.x, px), vx, __fmul_rn(d.x, px), y, H);
            const float dvy_dy = grad1(__fmul_rn(u.y, py), vy, __fmul_rn(d.y, py), y, H);
            const float div = __fadd_rn(dvx_dx, dvy_dy);
            const float curl = __fsub_rn(dvy_dx, dvx_dy);
            const int m = (fabsf(div) <= alpha) && (fabsf(curl) <= alpha);
            // v * mask as numpy computes it (main.py:600-601): a masked negative velocity is -0.0, a NaN stays NaN
            const float mf = m ? 1.f : 0.f;
            const float vxf = __fmul_rn(vx, mf), vyf = __fmul_rn(vy, mf);
            // sqrt is monotone and correctly rounded, so "sqrt(s) > thresh" is "s > s_crit" with s_crit the
            // largest double whose root is still <= thresh (found on the host): no DSQRT per cell.  The
            // f32 sum of squares (relative error < 2e-7) settles every cell that is not within 1e-6 of
            // the threshold; only those take the exact fp64 path.
            const float s32 = vxf * vxf + vyf * vyf;
            int is_valid;
            if (s32 > s_hi)
                is_valid = 1;
            else if (s32 < s_lo)
                is_valid = 0;
            else {
                const double dx = vxf, dy = vyf;
                is_valid = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) > s_crit;
            }
            n_ok += is_valid;
            if (vx_o) vx_o[base + o] = vx;
            if (vy_o) vy_o[base + o] = vy;
            if (ang_o) ang_o[base + o] = curl;
            if (mask_o) mask_o[base + o] = static_cast<uint8_t>(m);
            if (vxf_o) vxf_o[base + o] = vxf;
            if (vyf_o) vyf_o[base + o] = vyf;
            if (valid_o) valid_o[base + o] = static_cast<uint8_t>(is_valid);
        }
    }
    if (n_valid) {
        // warp reduction, then one atomic per warp
        n_ok = __reduce_add_sync(0xffffffffu, n_ok);
        if ((threadIdx.x & 31) == 0 && n_ok) atomicAdd(n_valid + b, n_ok);
    }
}

// The same for four consecutive columns of one row per thread (W % 4 == 0, 16-byte aligned arrays):
// 16-byte loads of the three flow rows and 16-byte / 4-byte stores of the outputs — a third of the
// load / store instructions of the column form, which is what bounds this stage.
__global__ void __launch_bounds__(256) k_velmask4(const float2* __restrict__ flow, int H, int W, float px, float py,
                                                  float alpha, double s_crit, float s_lo, float s_hi,
                                                  float* __restrict__ vx_o, float* __restrict__ vy_o,
                                                  float* __restrict__ ang_o, uint8_t* __restrict__ mask_o,
                                                  float* __restrict__ vxf_o, float* __restrict__ vyf_o,
                                                  uint8_t* __restrict__ valid_o, int32_t* __restrict__ n_valid) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y, b = blockIdx.z;
    int n_ok = 0;
    if (x4 < W) {
        const size_t base = static_cast<size_t>(b) * H * W;
        const float2* f = flow + base;
        const size_t o = static_cast<size_t>(y) * W + x4;
        const size_t ou = y > 0 ? o - W : o, od = y < H - 1 ? o + W : o;
        float2 c[6], u[4], d[4];   // c[0] / c[5]: the neighbours left and right of the four cells
        {
            const float4 a0 = *reinterpret_cast<const float4*>(f + o), a1 = *reinterpret_cast<const float4*>(f + o + 2);
            const float4 u0 = *reinterpret_cast<const float4*>(f + ou), u1 = *reinterpret_cast<const float4*>(f + ou + 2);
            const float4 d0 = *reinterpret_cast<const float4*>(f + od), d1 = *reinterpret_cast<const float4*>(f + od + 2);
            c[0] = f[o - (x4 > 0)];
            c[5] = f[o + 3 + (x4 + 4 < W)];
            c[1] = make_float2(a0.x, a0.y), c[2] = make_float2(a0.z, a0.w);
            c[3] = make_float2(a1.x, a1.y), c[4] = make_float2(a1.z, a1.w);
            u[0] = make_float2(u0.x, u0.y), u[1] = make_float2(u0.z, u0.w);
            u[2] = make_float2(u1.x, u1.y), u[3] = make_float2(u1.z, u1.w);
            d[0] = make_float2(d0.x, d0.y), d[1] = make_float2(d0.z, d0.w);
            d[2] = make_float2(d1.x, d1.y), d[3] = make_float2(d1.z, d1.w);
        }
        float vx[4], vy[4], cu[4], vxf[4], vyf[4];
        uint8_t mk[4], ok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x4 + j;
            const float2 cc = c[j + 1], l = c[j], r = c[j + 2];
            vx[j] = __fmul_rn(cc.x, px), vy[j] = __fmul_rn(cc.y, py);
            const float dvx_dx = grad1(__fmul_rn(l.x, px), vx[j], __fmul_rn(r.x, px), x, W);
            const float dvy_dx = grad1(__fmul_rn(l.y, py), vy[j], __fmul_rn(r.y, py), x, W);
            const float dvx_dy = grad1(__fmul_rn(u[j].x, px), vx[j], __fmul_rn(d[j].x, px), y, H);
            const float dvy_dy = grad1(__fmul_rn(u[j].y, py), vy[j], __fmul_rn(d[j].y, py), y, H);
            const float div = __fadd_rn(dvx_dx, dvy_dy);
            cu[j] = __fsub_rn(dvy_dx, dvx_dy);
            const int m = (fabsf(div) <= alpha) && (fabsf(cu[j]) <= alpha);
            const float mf = m ? 1.f : 0.f;   // v * mask like numpy: -0.0 for a masked negative velocity, NaN stays NaN
            vxf[j] = __fmul_rn(vx[j], mf), vyf[j] = __fmul_rn(vy[j], mf);
            const float s32 = vxf[j] * vxf[j] + vyf[j] * vyf[j];   // see k_velmask for the threshold logic
            int is_valid;
            if (s32 > s_hi)
                is_valid = 1;
            else if (s32 < s_lo)
                is_valid = 0;
            else {
                const double dx = vxf[j], dy = vyf[j];
                is_valid = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) > s_crit;
            }
            mk[j] = static_cast<uint8_t>(m), ok[j] = static_cast<uint8_t>(is_valid);
            n_ok += is_valid;
        }
        if (vx_o) *reinterpret_cast<float4*>(vx_o + base + o) = make_float4(vx[0], vx[1], vx[2], vx[3]);
        if (vy_o) *reinterpret_cast<float4*>(vy_o + base + o) = make_float4(vy[0], vy[1], vy[2], vy[3]);
        if (ang_o) *reinterpret_cast<float4*>(ang_o + base + o) = make_float4(cu[0], cu[1], cu[2], cu[3]);
        if (mask_o) *reinterpret_cast<uchar4*>(mask_o + base + o) = make_uchar4(mk[0], mk[1], mk[2], mk[3]);
        if (vxf_o) *reinterpret_cast<float4*>(vxf_o + base + o) = make_float4(vxf[0], vxf[1], vxf[2], vxf[3]);
        if (vyf_o) *reinterpret_cast<float4*>(vyf_o + base + o) = make_float4(vyf[0], vyf[1], vyf[2], vyf[3]);
        if (valid_o) *reinterpret_cast<uchar4*>(valid_o + base + o) = make_uchar4(ok[0], ok[1], ok[2], ok[3]);
    }
    if (n_valid) {
        n_ok = __reduce_add_sync(0xffffffffu, n_ok);
        if ((threadIdx.x & 31) == 0 && n_ok) atomicAdd(n_valid + b, n_ok);
    }
}

// curl of the FILTERED field (main.py:604-606); f64 arithmetic on f32-representable
// values, stored as f32 (the reference only writes it to a CSV).
__global__ void __launch_bounds__(256) k_curl_filtered(const float* __restrict__ vxf, const float* __restrict__ vyf,
                                                       int H, int W, float* __restrict__ ang_f) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const size_t base = static_cast<size_t>(b) * H * W;
    const size_t o = static_cast<size_t>(y) * W + x;
    const float* vx = vxf + base;
    const float* vy = vyf + base;
    double dvy_dx, dvx_dy;
    if (x == 0)
        dvy_dx = static_cast<double>(vy[o + 1]) - vy[o];
    else if (x == W - 1)
        dvy_dx = static_cast<double>(vy[o]) - vy[o - 1];
    else
        dvy_dx = (static_cast<double>(vy[o + 1]) - vy[o - 1]) / 2.0;
    if (y == 0)
        dvx_dy = static_cast<double>(vx[o + W]) - vx[o];
    else if (y == H - 1)
        dvx_dy = static_cast<double>(vx[o]) - vx[o - W];
    else
        dvx_dy = (static_cast<double>(vx[o + W]) - vx[o - W]) / 2.0;
    ang_f[base + o] = static_cast<float>(dvy_dx - dvx_dy);
}

// The filtered field as the reference holds it (main.py:600-606): v * mask is float64 there (f32 * int64),
// so are its magnitude and its curl (np.gradient of f64 arrays) — what the per-cell CSV and the .npy grids
// hold.  One pass from the f32 filtered velocities; any output may be NULL.
__global__ void __launch_bounds__(256) k_filtered_f64(const float* __restrict__ vxf, const float* __restrict__ vyf,
                                                      int H, int W, double* __restrict__ vx64,
                                                      double* __restrict__ vy64, double* __restrict__ mag64,
                                                      double* __restrict__ ang64) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const size_t base = static_cast<size_t>(b) * H * W;
    const size_t o = static_cast<size_t>(y) * W + x;
    const float* vx = vxf + base;
    const float* vy = vyf + base;
    const double dx = vx[o], dy = vy[o];
    if (vx64) vx64[base + o] = dx;
    if (vy64) vy64[base + o] = dy;
    if (mag64) mag64[base + o] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    if (ang64) {
        double dvy_dx, dvx_dy;
        if (x == 0)
            dvy_dx = static_cast<double>(vy[o + 1]) - vy[o];
        else if (x == W - 1)
            dvy_dx = static_cast<double>(vy[o]) - vy[o - 1];
        else
            dvy_dx = (static_cast<double>(vy[o + 1]) - vy[o - 1]) / 2.0;
        if (y == 0)
            dvx_dy = static_cast<double>(vx[o + W]) - vx[o];
        else if (y == H - 1)
            dvx_dy = static_cast<double>(vx[o]) - vx[o - W];
        else
            dvx_dy = (static_cast<double>(vx[o + W]) - vx[o - W]) / 2.0;
        ang64[base + o] = dvy_dx - dvx_dy;
    }
}

// f64 -> f32 with a check that nothing is lost (the reference's filtered velocities are f32 values held in
// f64, main.py:600-601); *flag is raised when some value is not f32-representable (NaN counts as fine)
__global__ void __launch_bounds__(256) k_narrow_checked(const double* __restrict__ src, size_t n,
                                                        float* __restrict__ dst, int* __restrict__ flag) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = src[i];
    const float f = static_cast<float>(v);
    dst[i] = f;
    if (static_cast<double>(f) != v && v == v) *flag = 1;
}

// ---- propagation masks (main.py:166-221) ---------------------------------------------------------
// The reference scatters every cell's velocity to the cell it would reach after dt, in row-major
// order, so the LAST source wins a contested target; then compares the scattered field with the
// actual one.  Here: atomicMax of the source rank per target, then a gather.  T is the dtype the
// reference would compute in (f32 for the velocities of compute_velocity_vectors, f64 otherwise);
// every operation is rounded separately, in numpy's order.
template <typename T>
struct PropOps;
template <>
struct PropOps<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <>
struct PropOps<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

template <typename T>
__global__ void __launch_bounds__(256) k_prop_scatter(const T* __restrict__ vx, const T* __restrict__ vy,
                                                      const T* __restrict__ ax, const T* __restrict__ ay, int H,
                                                      int W, T dt, T dt2, T gx, T gy, int32_t* __restrict__ winner) {
    using O = PropOps<T>;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, b = blockIdx.z;
    if (j >= W) return;
    const size_t base = static_cast<size_t>(b) * H * W;
    const int o = i * W + j;
    T sx = O::mul(vx[base + o], dt), sy = O::mul(vy[base + o], dt);
    if (ax) {
        sx = O::add(sx, O::mul(O::mul(static_cast<T>(0.5), ax[base + o]), dt2));
        sy = O::add(sy, O::mul(O::mul(static_cast<T>(0.5), ay[base + o]), dt2));
    }
    const T ti = O::add(static_cast<T>(i), floor(O::div(sx, gx)));
    const T tj = O::add(static_cast<T>(j), floor(O::div(sy, gy)));
    // int() truncates; NaN / inf (the reference raises on them) and out-of-grid targets do not propagate
    if (!(ti > static_cast<T>(-1) && ti < static_cast<T>(H) && tj > static_cast<T>(-1) && tj < static_cast<T>(W))) return;
    const int ii = static_cast<int>(ti), jj = static_cast<int>(tj);
    if (ii < 0 || jj < 0) return;  // (-1, 0) truncates to 0 and IS a hit; below -1 never gets here
    atomicMax(winner + base + static_cast<size_t>(ii) * W + jj, o);
}

template <typename T>
__global__ void __launch_bounds__(256) k_prop_mask(const T* __restrict__ vx, const T* __restrict__ vy, int H, int W,
                                                   T alpha, const int32_t* __restrict__ winner,
                                                   uint8_t* __restrict__ mask) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, b = blockIdx.z;
    if (j >= W) return;
    const size_t base = static_cast<size_t>(b) * H * W;
    const size_t o = base + static_cast<size_t>(i) * W + j;
    const int wn = winner[o];
    const T pvx = wn >= 0 ? vx[base + wn] : static_cast<T>(0), pvy = wn >= 0 ? vy[base + wn] : static_cast<T>(0);
    mask[o] = (fabs(pvx - vx[o]) <= alpha) && (fabs(pvy - vy[o]) <= alpha);
}

template <typename T>
int prop_run(datmo_ctx* h, const T* vx, const T* vy, const T* ax, const T* ay, int H, int W, int batch, double dt,
             double gx, double gy, double alpha_p, uint8_t* mask) {
    const size_t n = static_cast<size_t>(batch) * H * W;
    DATMO_TRY(datmo_ws_reserve(h, n * sizeof(int32_t) + 256));
    int32_t* winner = reinterpret_cast<int32_t*>(h->ws);
    DATMO_CHECK_CUDA(h, cudaMemsetAsync(winner, 0xff, n * sizeof(int32_t), h->stream));
    dim3 g(ceil_div(W, 256), H, batch);
    {
        LaunchScope ls(h, DATMO_TAG_VELMASK);
        k_prop_scatter<T><<<g, 256, 0, h->stream>>>(vx, vy, ax, ay, H, W, static_cast<T>(dt), static_cast<T>(dt * dt),
                                                    static_cast<T>(gx), static_cast<T>(gy), winner);
    }
    DATMO_POST_LAUNCH(h);
    {
        LaunchScope ls(h, DATMO_TAG_VELMASK);
        k_prop_mask<T><<<g, 256, 0, h->stream>>>(vx, vy, H, W, static_cast<T>(alpha_p), winner, mask);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

}  // namespace

extern "C" int datmo_propagation_mask_dev(datmo_handle_t h, const void* vx, const void* vy, const void* ax,
                                          const void* ay, int dtype, int H, int W, int batch, double dt, double grid_x,
                                          double grid_y, double alpha_p, uint8_t* mask) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, vx && vy && mask && H >= 1 && W >= 1 && batch >= 1, "bad arguments");
    DATMO_REQUIRE(h, (ax == nullptr) == (ay == nullptr), "ax and ay come together");
    DATMO_REQUIRE(h, dtype == DATMO_F32 || dtype == DATMO_F64, "dtype must be DATMO_F32 or DATMO_F64");
    DATMO_REQUIRE(h, static_cast<int64_t>(H) * W < (int64_t(1) << 31), "grid too large for int32 cell indices");
    DATMO_REQUIRE(h, H <= 65535 && batch <= 65535, "H and batch must fit a CUDA grid dimension");
    if (dtype == DATMO_F32)
        return prop_run<float>(h, static_cast<const float*>(vx), static_cast<const float*>(vy),
                               static_cast<const float*>(ax), static_cast<const float*>(ay), H, W, batch, dt, grid_x,
                               grid_y, alpha_p, mask);
    return prop_run<double>(h, static_cast<const double*>(vx), static_cast<const double*>(vy),
                            static_cast<const double*>(ax), static_cast<const double*>(ay), H, W, batch, dt, grid_x,
                            grid_y, alpha_p, mask);
}

extern "C" int datmo_velocity_mask_dev(datmo_handle_t h, const float* flow, int H, int W, int batch, double px_x,
                                       double px_y, double alpha_cont, double thresh, float* vx, float* vy, float* ang,
                                       uint8_t* mask, float* vx_f, float* vy_f, float* ang_f, uint8_t* valid,
                                       int32_t* n_valid) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, flow && H >= 2 && W >= 2 && batch >= 1, "need flow, H, W >= 2 (np.gradient) and batch >= 1");
    DATMO_REQUIRE(h, H <= 65535 && batch <= 65535, "H and batch must fit a CUDA grid dimension");
    DATMO_REQUIRE(h, !ang_f || (vx_f && vy_f), "ang_f needs vx_f and vy_f");
    if (n_valid) DATMO_CHECK_CUDA(h, cudaMemsetAsync(n_valid, 0, batch * sizeof(int32_t), h->stream));
    // largest s with sqrt(s) <= thresh
    double s_crit = -1.0;
    if (thresh >= 0) {
        s_crit = thresh * thresh;
        while (s_crit > 0 && sqrt(s_crit) > thresh) s_crit = nextafter(s_crit, 0.0);
        while (sqrt(nextafter(s_crit, INFINITY)) <= thresh) s_crit = nextafter(s_crit, INFINITY);
    } else if (thresh != thresh) {
        s_crit = INFINITY;  // NaN threshold: nothing is valid
    }
    // f32 guard band around s_crit; outside it the f32 sum of squares decides
    float s_lo = -INFINITY, s_hi = INFINITY;  // degenerate thresholds: every cell takes the exact path
    if (s_crit > 1e-30 && s_crit < 1e30) {
        s_lo = static_cast<float>(s_crit * (1.0 - 1e-6));
        s_hi = static_cast<float>(s_crit * (1.0 + 1e-6));
    }
    dim3 g(ceil_div(W, 256), H, batch);
    dim3 gv(ceil_div(W, 256), ceil_div(H, VM_ROWS), batch);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = (W & 3) == 0 && al16(flow) && al16(vx) && al16(vy) && al16(ang) && al16(vx_f) && al16(vy_f) &&
                     (reinterpret_cast<uintptr_t>(mask) & 3) == 0 && (reinterpret_cast<uintptr_t>(valid) & 3) == 0;
    {
        LaunchScope ls(h, DATMO_TAG_VELMASK);
        if (vec) {
            dim3 g4(ceil_div(W, 1024), H, batch);
            k_velmask4<<<g4, 256, 0, h->stream>>>(reinterpret_cast<const float2*>(flow), H, W,
                                                  static_cast<float>(px_x), static_cast<float>(px_y),
                                                  static_cast<float>(alpha_cont), s_crit, s_lo, s_hi, vx, vy, ang, mask,
                                                  vx_f, vy_f, valid, n_valid);
        } else {
            k_velmask<<<gv, 256, 0, h->stream>>>(reinterpret_cast<const float2*>(flow), H, W, static_cast<float>(px_x),
                                                 static_cast<float>(px_y), static_cast<float>(alpha_cont), s_crit, s_lo,
                                                 s_hi, vx, vy, ang, mask, vx_f, vy_f, valid, n_valid);
        }
    }
    DATMO_POST_LAUNCH(h);
    if (ang_f) {
        LaunchScope ls(h, DATMO_TAG_VELMASK);
        k_curl_filtered<<<g, 256, 0, h->stream>>>(vx_f, vy_f, H, W, ang_f);
        DATMO_POST_LAUNCH(h);
    }
    return DATMO_OK;
}

extern "C" int datmo_filtered_grids_f64_dev(datmo_handle_t h, const float* vx_f, const float* vy_f, int H, int W,
                                            int batch, double* vx64, double* vy64, double* mag64, double* ang64) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, vx_f && vy_f && H >= 2 && W >= 2 && batch >= 1, "need vx_f, vy_f, H, W >= 2 and batch >= 1");
    DATMO_REQUIRE(h, H <= 65535 && batch <= 65535, "H and batch must fit a CUDA grid dimension");
    dim3 g(ceil_div(W, 256), H, batch);
    {
        LaunchScope ls(h, DATMO_TAG_VELMASK);
        k_filtered_f64<<<g, 256, 0, h->stream>>>(vx_f, vy_f, H, W, vx64, vy64, mag64, ang64);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

extern "C" int datmo_narrow_f64_dev(datmo_handle_t h, const double* src, int64_t n, float* dst, int* lossy) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, src && dst && lossy && n >= 1, "bad arguments");
    DATMO_TRY(datmo_ws_reserve(h, 256));
    int* d_flag = reinterpret_cast<int*>(h->ws);
    DATMO_CHECK_CUDA(h, cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
    {
        LaunchScope ls(h, DATMO_TAG_VELMASK);
        k_narrow_checked<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, h->stream>>>(src, static_cast<size_t>(n), dst, d_flag);
    }
    DATMO_POST_LAUNCH(h);
    DATMO_CHECK_CUDA(h, cudaMemcpyAsync(lossy, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    return DATMO_OK;
}
