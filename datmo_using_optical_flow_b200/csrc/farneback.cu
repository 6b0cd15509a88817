// Dense Farneback optical flow for batches of BEV frame pairs, sm_100a.
//
// Replaces cv2.calcOpticalFlowFarneback as the reference calls it at
// Optical_flow/main.py:131-142 (flags = 0: box window, no initial flow).  The
// algorithm (OpenCV video/optflowgf.cpp) is re-derived for
// the GPU, not transcribed:
//   * the pyramid image of a layer is evaluated directly at the resized resolution:
//     the separable Gaussian is applied only at the source taps the bilinear resize
//     reads (k_pyr_h, k_pyr_v), instead of blurring the full-resolution frame;
//   * polynomial expansion is one tiled kernel, both separable passes staged through
//     shared memory (k_polyexp), for prev and next frames of the whole batch at once;
//   * the finest layer's pyramid image + polynomial expansion is ONE kernel straight from the
//     uint8 frame (k_pyr0_polyexp);
//   * one flow iteration = updateMatrices + 5-channel box blur + 2x2 solve is ONE
//     kernel (k_flow_iter_xm for the reference's window, k_flow_iter otherwise): the M field
//     never exists in global memory;
//   * the 5-coefficient fields R0 / R1 are stored per array of B images as
//     [B][h*w] float4 (coefficients 0..3) followed by [B][h*w] float (coefficient 4): a
//     bilinear tap is one 16-byte + one 4-byte load instead of five scalar ones, and a warp
//     still reads whole 128-byte lines; flow is float2 per pixel; M (unfused path) is planar.
// Arithmetic is f32 with direct (non-running) window sums, which SURVEY.md §3.2 F6
// and tests/test_oracle_farneback.py show stays inside the parity tolerance.
#include <math.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------
// host-side plan (F0 / F1)
// ------------------------------------------------------------------------------------
struct FbLayer {
    int k;
    double scale, sigma;
    int ksize, w, h;
};

inline int cv_round(double v) { return static_cast<int>(nearbyint(v)); }  // round half to even

std::vector<FbLayer> fb_plan(int H, int W, double pyr_scale, int levels) {
    const int min_size = 32;
    int k = 0;
    double scale = 1.0;
    while (k < levels) {
        scale *= pyr_scale;
        if (W * scale < min_size || H * scale < min_size) break;
        ++k;
    }
    std::vector<FbLayer> out;
    for (int kk = k; kk >= 0; --kk) {
        double s = 1.0;
        for (int i = 0; i < kk; ++i) s *= pyr_scale;
        FbLayer L;
        L.k = kk;
        L.scale = s;
        L.sigma = (1.0 / s - 1.0) * 0.5;
        L.ksize = std::max(cv_round(L.sigma * 5) | 1, 3);
        L.w = cv_round(W * s);
        L.h = cv_round(H * s);
        out.push_back(L);
    }
    return out;
}

std::vector<float> gaussian_kernel(int ksize, double sigma) {
    std::vector<float> k(ksize);
    if (sigma <= 0 && ksize == 3) {
        k[0] = 0.25f, k[1] = 0.5f, k[2] = 0.25f;
        return k;
    }
    if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    std::vector<double> t(ksize);
    double sum = 0;
    for (int i = 0; i < ksize; ++i) {
        double x = i - (ksize - 1) * 0.5;
        t[i] = exp(-0.5 / (sigma * sigma) * x * x);
        sum += t[i];
    }
    for (int i = 0; i < ksize; ++i) k[i] = static_cast<float>(t[i] / sum);
    return k;
}

// INTER_LINEAR source tap of destination index d (F2), host side: src = (d+0.5)*S/D - 0.5 in fp64,
// then the two clamp rules.
inline void resize_tap_host(int d, int S, double ratio, int& s, double& f) {
    const double src = (d + 0.5) * ratio - 0.5;
    const double fl = floor(src);
    s = static_cast<int>(fl);
    f = src - fl;
    if (s < 0) s = 0, f = 0;
    if (s >= S - 1) s = S - 1, f = 0;
}

// Tables of one pyramid layer, appended to `out` as 32-bit words (the block starts 8-byte aligned):
//   hf [w] double | vf [h] double | hx0 [w] int | vy0 [h] int | g [ksize] float (+ pad to even)
// hx0 = first source column of an output column's blur window (tap - ksize / 2), hf its blend
// weight towards the next tap; likewise for rows.
size_t pyr_table_words(const FbLayer& L) {
    return 3 * static_cast<size_t>(L.w + L.h) + ((L.ksize + 1) & ~1) + ((L.w + L.h) & 1);
}

void pyr_tables(int H, int W, const FbLayer& L, std::vector<float>& out) {
    const int r = L.ksize >> 1;
    const std::vector<float> g = gaussian_kernel(L.ksize, L.sigma);
    const size_t base = out.size();  // even by construction
    out.resize(base + pyr_table_words(L), 0.f);
    char* p = reinterpret_cast<char*>(out.data() + base);
    char* hf = p;
    char* vf = hf + sizeof(double) * L.w;
    char* hx = vf + sizeof(double) * L.h;
    char* vy = hx + sizeof(int32_t) * L.w;
    char* gk = vy + sizeof(int32_t) * L.h;
    auto fill = [&](int n_out, int n_src, char* first, char* frac) {
        const double ratio = static_cast<double>(n_src) / n_out;
        for (int d = 0; d < n_out; ++d) {
            int sidx;
            double f;
            resize_tap_host(d, n_src, ratio, sidx, f);
            const int32_t x0 = sidx - r;
            memcpy(first + sizeof(int32_t) * d, &x0, sizeof(int32_t));
            memcpy(frac + sizeof(double) * d, &f, sizeof(double));
        }
    };
    fill(L.w, W, hx, hf);
    fill(L.h, H, vy, vf);
    memcpy(gk, g.data(), sizeof(float) * L.ksize);
}

// Tables of one flow upsampling step (wi x hi -> wo x ho), appended to `out` as 32-bit words:
//   sx [wo] int | fx [wo] float | sy [ho] int | fy [ho] float
size_t up_table_words(int wo, int ho) { return 2 * static_cast<size_t>(wo + ho); }

void up_tables(int hi, int wi, int ho, int wo, std::vector<float>& out) {
    const size_t base = out.size();
    out.resize(base + up_table_words(wo, ho), 0.f);
    float* sx = out.data() + base;
    float* fx = sx + wo;
    float* sy = fx + wo;
    float* fy = sy + ho;
    auto fill = [&](int n_out, int n_src, float* first, float* frac) {
        const double ratio = static_cast<double>(n_src) / n_out;
        for (int d = 0; d < n_out; ++d) {
            int sidx;
            double f;
            resize_tap_host(d, n_src, ratio, sidx, f);
            const int32_t i32 = sidx;
            memcpy(first + d, &i32, sizeof(int32_t));
            frac[d] = static_cast<float>(f);
        }
    };
    fill(wo, wi, sx, fx);
    fill(ho, hi, sy, fy);
}

constexpr int POLY_MAX_N = 16;
struct PolyCoef {
    float g[POLY_MAX_N + 1], xg[POLY_MAX_N + 1], xxg[POLY_MAX_N + 1];
    float ig11, ig03, ig33, ig55;
    int n;
};

bool invert6(double a[6][6], double inv[6][6]) {
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) inv[i][j] = i == j;
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r)
            if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        if (a[piv][c] == 0) return false;
        for (int j = 0; j < 6; ++j) {
            std::swap(a[c][j], a[piv][j]);
            std::swap(inv[c][j], inv[piv][j]);
        }
        double d = 1.0 / a[c][c];
        for (int j = 0; j < 6; ++j) {
            a[c][j] *= d;
            inv[c][j] *= d;
        }
        for (int r = 0; r < 6; ++r) {
            if (r == c) continue;
            double f = a[r][c];
            if (f == 0) continue;
            for (int j = 0; j < 6; ++j) {
                a[r][j] -= f * a[c][j];
                inv[r][j] -= f * inv[c][j];
            }
        }
    }
    return true;
}

bool poly_setup(int n, double sigma, PolyCoef& pc) {
    if (n < 1 || n > POLY_MAX_N) return false;
    if (sigma < 1.1920928955078125e-07) sigma = n * 0.3;
    std::vector<float> g(2 * n + 1), xg(2 * n + 1), xxg(2 * n + 1);
    double s = 0;
    for (int x = -n; x <= n; ++x) {
        g[x + n] = static_cast<float>(exp(-x * x / (2 * sigma * sigma)));
        s += g[x + n];
    }
    s = 1.0 / s;
    for (int x = -n; x <= n; ++x) {
        g[x + n] = static_cast<float>(g[x + n] * s);
        xg[x + n] = static_cast<float>(x * g[x + n]);
        xxg[x + n] = static_cast<float>(x * x * g[x + n]);
    }
    double G[6][6] = {{0}};
    for (int y = -n; y <= n; ++y)
        for (int x = -n; x <= n; ++x) {
            float gg = g[y + n] * g[x + n];
            G[0][0] += gg;
            G[1][1] += gg * x * x;
            G[3][3] += gg * x * x * x * x;
            G[5][5] += gg * x * x * y * y;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    double inv[6][6];
    if (!invert6(G, inv)) return false;
    pc.n = n;
    for (int k = 0; k <= n; ++k) {
        pc.g[k] = g[n + k];
        pc.xg[k] = xg[n + k];
        pc.xxg[k] = xxg[n + k];
    }
    pc.ig11 = static_cast<float>(inv[1][1]);
    pc.ig03 = static_cast<float>(inv[0][3]);
    pc.ig33 = static_cast<float>(inv[3][3]);
    pc.ig55 = static_cast<float>(inv[5][5]);
    return true;
}

// ------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

__device__ __forceinline__ float load_px(const uint8_t* p) { return static_cast<float>(*p); }
__device__ __forceinline__ float load_px(const float* p) { return *p; }

// View of image b inside an R array of B images: [B][n] float4 | [B][n] float.
struct RView {
    const float4* q;
    const float* s;
};
__device__ __forceinline__ RView r_view(const float* R, int B, int b, size_t n) {
    RView v;
    v.q = reinterpret_cast<const float4*>(R) + static_cast<size_t>(b) * n;
    v.s = R + 4 * static_cast<size_t>(B) * n + static_cast<size_t>(b) * n;
    return v;
}
// floats from one R array of B images to the next (kept 16-byte aligned for the float4 part)
__host__ __device__ __forceinline__ size_t r_array_stride(int B, size_t n) {
    return (5 * static_cast<size_t>(B) * n + 63) & ~static_cast<size_t>(63);
}
// image index bz of a launch over several consecutive R arrays of B images each
__device__ __forceinline__ void r_out(float* R, int B, int bz, size_t n, float4*& q, float*& sdst) {
    const int arr = bz < B ? 0 : (bz < 2 * B ? 1 : bz / B), b = bz - arr * B;   // one or two arrays: no division
    float* base = R + static_cast<size_t>(arr) * r_array_stride(B, n);
    q = reinterpret_cast<float4*>(base) + static_cast<size_t>(b) * n;
    sdst = base + 4 * static_cast<size_t>(B) * n + static_cast<size_t>(b) * n;
}

// ------------------------------------------------------------------------------------
// F2: pyramid image.  Horizontal Gaussian + horizontal resize, then vertical Gaussian
// + vertical resize; the blur is only evaluated at the taps the resize reads.
// ------------------------------------------------------------------------------------
// The source tap and the fp64 blend weight of every output column (row) do not depend on the row
// (column): the host tabulates them once per layer (pyr_tables).
// Horizontal pass: a CTA stages PYR_ROWS source rows in shared memory as f32 (BORDER_REFLECT_101
// applied while staging, 16 pixels per load) and every thread produces one output column for all
// of them.  Arithmetic (order of the f32 tap sums, fp64 blend) is that of cv2's blur-then-resize.
constexpr int PYR_ROWS = 8;    // source rows staged by one CTA of the horizontal pass
constexpr int PYR_VROWS = 4;   // output rows per thread of the vertical pass

// four pixels -> one 16-byte shared store (consecutive lanes, consecutive chunks: conflict-free)
__device__ __forceinline__ float4 load4_px(const uint8_t* p) {
    const unsigned v = *reinterpret_cast<const unsigned*>(p);
    return make_float4(static_cast<float>(v & 0xffu), static_cast<float>((v >> 8) & 0xffu),
                       static_cast<float>((v >> 16) & 0xffu), static_cast<float>(v >> 24));
}
__device__ __forceinline__ float4 load4_px(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <typename SrcT>
__global__ void __launch_bounds__(256) k_pyr_h(const SrcT* __restrict__ src0, const SrcT* __restrict__ src1,
                                               int n0, float* __restrict__ T, int H, int W, int w,
                                               const float* __restrict__ kern, int ksize,
                                               const int* __restrict__ x0tab, const double* __restrict__ ftab,
                                               int vec) {
    extern __shared__ __align__(16) float srow[];  // [PYR_ROWS][SW]; element LM + c of a row is source column c
    const int r = ksize >> 1, LM = (r + 3) & ~3, SW = LM + ((W + r + 1 + 3) & ~3);
    const int y0 = blockIdx.x * PYR_ROWS, b = blockIdx.y;
    // images 0 .. n0-1 come from src0, the rest from src1 (prev and next frames in one launch)
    const SrcT* src = b < n0 ? src0 : src1 - static_cast<size_t>(n0) * H * W;
    if (vec) {
        // interior, four pixels per thread and trip (W % 4 == 0 and a suitably aligned image)
        const int per_row = W >> 2;
#pragma unroll 4
        for (int i = threadIdx.x; i < PYR_ROWS * per_row; i += 256) {
            const int rr = i / per_row, j = i - rr * per_row;
            const int y = min(y0 + rr, H - 1);
            *reinterpret_cast<float4*>(srow + rr * SW + LM + 4 * j) =
                load4_px(src + (static_cast<size_t>(b) * H + y) * W + 4 * j);
        }
        // the reflected margins: r columns on the left, r + 1 on the right
        const int nm = 2 * r + 1;
        for (int i = threadIdx.x; i < PYR_ROWS * nm; i += 256) {
            const int rr = i / nm, m = i - rr * nm;
            const int c = m < r ? m - r : W + (m - r);
            const int y = min(y0 + rr, H - 1);
            srow[rr * SW + LM + c] = load_px(src + (static_cast<size_t>(b) * H + y) * W + reflect101(c, W));
        }
    } else {
        const int span = W + 2 * r + 1;
        for (int i = threadIdx.x; i < PYR_ROWS * span; i += 256) {
            const int rr = i / span, c = i - rr * span - r;
            const int y = min(y0 + rr, H - 1);
            srow[rr * SW + LM + c] = load_px(src + (static_cast<size_t>(b) * H + y) * W + reflect101(c, W));
        }
    }
    __syncthreads();
    for (int xo = threadIdx.x; xo < w; xo += 256) {
        const float* s0 = srow + x0tab[xo] + LM;
        const double fx = ftab[xo];
        float a0[PYR_ROWS], a1[PYR_ROWS];
#pragma unroll
        for (int rr = 0; rr < PYR_ROWS; ++rr) a0[rr] = 0.f, a1[rr] = 0.f;
        for (int i = 0; i < ksize; ++i) {
            const float kv = kern[i];
#pragma unroll
            for (int rr = 0; rr < PYR_ROWS; ++rr) {
                a0[rr] = fmaf(kv, s0[rr * SW + i], a0[rr]);
                a1[rr] = fmaf(kv, s0[rr * SW + i + 1], a1[rr]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < PYR_ROWS; ++rr)
            if (y0 + rr < H)
                T[(static_cast<size_t>(b) * H + y0 + rr) * w + xo] =
                    fx == 0.0 ? a0[rr] : static_cast<float>((1.0 - fx) * a0[rr] + fx * a1[rr]);
    }
}

// Horizontal pass for uint8 frames with the lanes of a warp on 32 ROWS of one output column: the tap offset and
// the blend weight are warp-uniform, the staged rows stay bytes (row stride an odd number of 32-bit words, so the
// 32 lanes reading one column hit 32 banks — the column-per-lane kernel above strides its lanes by 1 / scale
// pixels and pays 3 - 4 wavefronts per load on the finer layers, and leaves most of its threads idle on the
// coarse ones), the taps are immediate operands, and the finished 32 x w block leaves through a transposing tile
// so the global stores are whole lines.  Same arithmetic, same T.
constexpr int PH2_ROWS = 32, PH2_THREADS = 512, PH2_MAX_K = 64;   // 16 warps share the staged rows
struct PyrTaps {
    float k[PH2_MAX_K];
};

// uint8 -> f32 on the integer and FMA pipes (I2F runs on the quarter-rate conversion unit, and a tap loop that
// converts every byte it reads is bound by it): 0x4B000000 | b is the float 2^23 + b, exactly
__device__ __forceinline__ float byte_to_float(unsigned b) { return __uint_as_float(0x4B000000u | b) - 8388608.f; }

template <int KS>   // compile-time tap count; 0 = runtime (ksize_rt)
__global__ void __launch_bounds__(PH2_THREADS) k_pyr_h_rows(const uint8_t* __restrict__ src0,
                                                            const uint8_t* __restrict__ src1, int n0,
                                                            float* __restrict__ T, int H, int W, int w, PyrTaps taps,
                                                            int ksize_rt, const int* __restrict__ x0tab,
                                                            const double* __restrict__ ftab, int vec) {
    extern __shared__ __align__(16) unsigned char sm8[];
    const int ksize = KS ? KS : ksize_rt;
    const int r = ksize >> 1, LM = (r + 3) & ~3;
    const int SWB = 4 * (((LM + W + r + 1 + 3) >> 2) | 1);   // bytes per staged row
    const int TS = w | 1;                                    // floats per row of the transposing tile
    unsigned char* sB = sm8;                                                   // [PH2_ROWS][SWB]
    float* sT = reinterpret_cast<float*>(sm8 + PH2_ROWS * SWB);                // [PH2_ROWS][TS]
    const int y0 = blockIdx.x * PH2_ROWS, b = blockIdx.y;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int NW = blockDim.x >> 5;
    const uint8_t* src = (b < n0 ? src0 : src1 - static_cast<size_t>(n0) * H * W) + static_cast<size_t>(b) * H * W;
    for (int rr = wi; rr < PH2_ROWS; rr += NW) {
        const uint8_t* grow = src + static_cast<size_t>(min(y0 + rr, H - 1)) * W;
        unsigned char* srow = sB + rr * SWB + LM;
        if (vec) {
            const int per_row = W >> 2;
#pragma unroll 8
            for (int j = lane; j < per_row; j += 32)
                reinterpret_cast<unsigned*>(srow)[j] = __ldg(reinterpret_cast<const unsigned*>(grow) + j);
            // the reflected margins: r columns on the left, r + 1 on the right
            for (int m = lane; m < 2 * r + 1; m += 32) {
                const int c = m < r ? m - r : W + (m - r);
                srow[c] = grow[reflect101(c, W)];
            }
        } else {
            for (int c = lane - r; c < W + r + 1; c += 32) srow[c] = grow[reflect101(c, W)];
        }
    }
    __syncthreads();
    const unsigned char* rowp = sB + lane * SWB + LM;
    float* trow = sT + lane * TS;
    int xo = wi;
    int x0n = xo < w ? x0tab[xo] : 0;
    double fxn = xo < w ? ftab[xo] : 0.0;
    for (; xo < w; xo += NW) {
        const unsigned char* s = rowp + x0n;
        const double fx = fxn;
        if (xo + NW < w) x0n = x0tab[xo + NW], fxn = ftab[xo + NW];   // the next column's table entries are in flight
        float a0 = 0.f, a1 = 0.f;
        if (KS) {
            float v = byte_to_float(s[0]);
            a0 = fmaf(taps.k[0], v, a0);
#pragma unroll
            for (int q = 1; q < (KS ? KS : 1); ++q) {
                v = byte_to_float(s[q]);
                a0 = fmaf(taps.k[q], v, a0);
                a1 = fmaf(taps.k[q - 1], v, a1);
            }
            v = byte_to_float(s[KS]);
            a1 = fmaf(taps.k[(KS ? KS : 1) - 1], v, a1);
        } else {
            float v = byte_to_float(s[0]);
            float kq = taps.k[0];
            a0 = fmaf(kq, v, a0);
            for (int q = 1; q < ksize; ++q) {
                const float kp = kq;
                kq = taps.k[q];
                v = byte_to_float(s[q]);
                a0 = fmaf(kq, v, a0);
                a1 = fmaf(kp, v, a1);
            }
            v = byte_to_float(s[ksize]);
            a1 = fmaf(kq, v, a1);
        }
        trow[xo] = fx == 0.0 ? a0 : static_cast<float>((1.0 - fx) * a0 + fx * a1);
    }
    __syncthreads();
    for (int rr = wi; rr < PH2_ROWS; rr += NW) {
        const int y = y0 + rr;
        if (y >= H) break;
        float* dst = T + (static_cast<size_t>(b) * H + y) * w;
        const float* t = sT + rr * TS;
        for (int x = lane; x < w; x += 32) dst[x] = t[x];
    }
}

// Vertical pass: a thread owns an output column and PYR_VROWS consecutive output rows whose tap
// chains are independent, so their loads overlap.
__global__ void __launch_bounds__(128) k_pyr_v(const float* __restrict__ T, float* __restrict__ out, int H, int w,
                                               int h, const float* __restrict__ kern, int ksize,
                                               const int* __restrict__ y0tab, const double* __restrict__ ftab) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    const int yo0 = blockIdx.y * PYR_VROWS, b = blockIdx.z;
    if (xo >= w) return;
    const float* base = T + static_cast<size_t>(b) * H * w + xo;
    int ys[PYR_VROWS];
    float a0[PYR_VROWS], a1[PYR_VROWS];
#pragma unroll
    for (int rr = 0; rr < PYR_VROWS; ++rr) {
        ys[rr] = y0tab[min(yo0 + rr, h - 1)];
        a0[rr] = 0.f, a1[rr] = 0.f;
    }
    for (int i = 0; i <= ksize; ++i) {
        // source row ys + i feeds tap i of a0 and tap i - 1 of a1: one load for both
        const float k0 = i < ksize ? kern[i] : 0.f, k1 = i > 0 ? kern[i - 1] : 0.f;
        float v[PYR_VROWS];
#pragma unroll
        for (int rr = 0; rr < PYR_VROWS; ++rr) {
            int y = ys[rr] + i;
            if (static_cast<unsigned>(y) >= static_cast<unsigned>(H)) y = reflect101(y, H);
            v[rr] = base[static_cast<size_t>(y) * w];
        }
#pragma unroll
        for (int rr = 0; rr < PYR_VROWS; ++rr) {
            if (i < ksize) a0[rr] = fmaf(k0, v[rr], a0[rr]);
            if (i > 0) a1[rr] = fmaf(k1, v[rr], a1[rr]);
        }
    }
#pragma unroll
    for (int rr = 0; rr < PYR_VROWS; ++rr) {
        const int yo = yo0 + rr;
        if (yo < h) {
            const double fy = ftab[yo];
            out[(static_cast<size_t>(b) * h + yo) * w + xo] =
                fy == 0.0 ? a0[rr] : static_cast<float>((1.0 - fy) * a0[rr] + fy * a1[rr]);
        }
    }
}

// The same pass with the taps as immediate operands, 32-bit in-image offsets and the border rows decided once per
// CTA (the four output rows of a CTA are the same for all its threads): k_pyr_v spends ~250 instructions per output
// on 64-bit address arithmetic, per-load reflection tests and tap loads.  Same arithmetic, same result.
template <int KS>   // compile-time tap count; 0 = runtime (ksize_rt)
__global__ void __launch_bounds__(128) k_pyr_v_taps(const float* __restrict__ T, float* __restrict__ out, int H, int w,
                                                    int h, PyrTaps taps, int ksize_rt,
                                                    const int* __restrict__ y0tab, const double* __restrict__ ftab) {
    const int ksize = KS ? KS : ksize_rt;
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    const int yo0 = blockIdx.y * PYR_VROWS, b = blockIdx.z;
    if (xo >= w) return;
    const float* base = T + static_cast<size_t>(b) * H * w + xo;
    int ys[PYR_VROWS];
#pragma unroll
    for (int rr = 0; rr < PYR_VROWS; ++rr) ys[rr] = __ldg(y0tab + min(yo0 + rr, h - 1));
    float a0[PYR_VROWS], a1[PYR_VROWS];
    if (ys[0] >= 0 && ys[PYR_VROWS - 1] + ksize < H) {   // the tables are monotone: every tap row is inside the image
        const float* p[PYR_VROWS];
#pragma unroll
        for (int rr = 0; rr < PYR_VROWS; ++rr) p[rr] = base + ys[rr] * w;
        if (KS) {
            float v[PYR_VROWS];
#pragma unroll
            for (int rr = 0; rr < PYR_VROWS; ++rr) v[rr] = p[rr][0];
#pragma unroll
            for (int rr = 0; rr < PYR_VROWS; ++rr) a0[rr] = fmaf(taps.k[0], v[rr], 0.f), a1[rr] = 0.f;
#pragma unroll
            for (int q = 1; q < (KS ? KS : 1); ++q) {
#pragma unroll
                for (int rr = 0; rr < PYR_VROWS; ++rr) v[rr] = p[rr][q * w];
#pragma unroll
                for (int rr = 0; rr < PYR_VROWS; ++rr) {
                    a0[rr] = fmaf(taps.k[q], v[rr], a0[rr]);
                    a1[rr] = fmaf(taps.k[q - 1], v[rr], a1[rr]);
                }
            }
#pragma unroll
            for (int rr = 0; rr < PYR_VROWS; ++rr) a1[rr] = fmaf(taps.k[(KS ? KS : 1) - 1], p[rr][KS * w], a1[rr]);
        } else {
#pragma unroll
            for (int rr = 0; rr < PYR_VROWS; ++rr) a0[rr] = 0.f, a1[rr] = 0.f;
            float kp = 0.f;
            for (int q = 0; q <= ksize; ++q) {
                const float kq = q < ksize ? taps.k[q] : 0.f;
#pragma unroll
                for (int rr = 0; rr < PYR_VROWS; ++rr) {
                    const float v = p[rr][q * w];
                    if (q < ksize) a0[rr] = fmaf(kq, v, a0[rr]);
                    if (q > 0) a1[rr] = fmaf(kp, v, a1[rr]);
                }
                kp = kq;
            }
        }
    } else {
#pragma unroll
        for (int rr = 0; rr < PYR_VROWS; ++rr) a0[rr] = 0.f, a1[rr] = 0.f;
        float kp = 0.f;
        for (int q = 0; q <= ksize; ++q) {
            const float kq = q < ksize ? taps.k[q] : 0.f;
#pragma unroll
            for (int rr = 0; rr < PYR_VROWS; ++rr) {
                const float v = base[reflect101(ys[rr] + q, H) * w];
                if (q < ksize) a0[rr] = fmaf(kq, v, a0[rr]);
                if (q > 0) a1[rr] = fmaf(kp, v, a1[rr]);
            }
            kp = kq;
        }
    }
#pragma unroll
    for (int rr = 0; rr < PYR_VROWS; ++rr) {
        const int yo = yo0 + rr;
        if (yo < h) {
            const double fy = __ldg(ftab + yo);
            out[(static_cast<size_t>(b) * h + yo) * w + xo] =
                fy == 0.0 ? a0[rr] : static_cast<float>((1.0 - fy) * a0[rr] + fy * a1[rr]);
        }
    }
}

// ------------------------------------------------------------------------------------
// F4: polynomial expansion, one tile per CTA, both passes through shared memory
// ------------------------------------------------------------------------------------
constexpr int PE_TX = 32, PE_TY = 16, PE_THREADS = 256;

__global__ void __launch_bounds__(PE_THREADS) k_polyexp(const float* __restrict__ I, float* __restrict__ R, int w,
                                                        int h, int imgs_per_array, PolyCoef pc) {
    extern __shared__ float smem[];
    const int n = pc.n;
    const int RW = PE_TX + 2 * n, RH = PE_TY + 2 * n;
    float* sI = smem;                  // [RH][RW]
    float* sr0 = sI + RH * RW;         // [PE_TY][RW] x3
    float* sr1 = sr0 + PE_TY * RW;
    float* sr2 = sr1 + PE_TY * RW;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * PE_TX, y0 = blockIdx.y * PE_TY, b = blockIdx.z;
    const size_t plane = static_cast<size_t>(w) * h;
    const float* Ib = I + b * plane;
    for (int i = tid; i < RH * RW; i += PE_THREADS) {
        int yy = i / RW, xx = i - yy * RW;
        int gx = min(max(x0 - n + xx, 0), w - 1);
        int gy = min(max(y0 - n + yy, 0), h - 1);
        sI[i] = Ib[static_cast<size_t>(gy) * w + gx];
    }
    __syncthreads();
    // vertical pass (rows replicate: the clamp above already did it)
    for (int i = tid; i < PE_TY * RW; i += PE_THREADS) {
        int ty = i / RW, xx = i - ty * RW;
        const float* c = sI + (ty + n) * RW + xx;
        float r0 = c[0] * pc.g[0], r1 = 0.f, r2 = 0.f;
        for (int k = 1; k <= n; ++k) {
            float s0 = c[-k * RW], s1 = c[k * RW];
            float p = s0 + s1;
            r0 = fmaf(pc.g[k], p, r0);
            r1 = fmaf(pc.xg[k], s1 - s0, r1);
            r2 = fmaf(pc.xxg[k], p, r2);
        }
        sr0[i] = r0;
        sr1[i] = r1;
        sr2[i] = r2;
    }
    __syncthreads();
    // horizontal pass
    float4* Rq;
    float* Rs;
    r_out(R, imgs_per_array, b, plane, Rq, Rs);
    for (int i = tid; i < PE_TX * PE_TY; i += PE_THREADS) {
        int ty = i / PE_TX, tx = i - ty * PE_TX;
        int gx = x0 + tx, gy = y0 + ty;
        if (gx >= w || gy >= h) continue;
        const float* p0 = sr0 + ty * RW + tx + n;
        const float* p1 = sr1 + ty * RW + tx + n;
        const float* p2 = sr2 + ty * RW + tx + n;
        float b1 = p0[0] * pc.g[0], b3 = p1[0] * pc.g[0], b5 = p2[0] * pc.g[0];
        float b2 = 0.f, b4 = 0.f, b6 = 0.f;
        for (int k = 1; k <= n; ++k) {
            float tg = p0[k] + p0[-k];
            b1 = fmaf(tg, pc.g[k], b1);
            b4 = fmaf(tg, pc.xxg[k], b4);
            b2 = fmaf(p0[k] - p0[-k], pc.xg[k], b2);
            b3 = fmaf(p1[k] + p1[-k], pc.g[k], b3);
            b6 = fmaf(p1[k] - p1[-k], pc.xg[k], b6);
            b5 = fmaf(p2[k] + p2[-k], pc.g[k], b5);
        }
        size_t o = static_cast<size_t>(gy) * w + gx;
        Rq[o] = make_float4(b3 * pc.ig11, b2 * pc.ig11, fmaf(b1, pc.ig03, b5 * pc.ig33),
                            fmaf(b1, pc.ig03, b4 * pc.ig33));
        Rs[o] = b6 * pc.ig55;
    }
}

// ------------------------------------------------------------------------------------
// F2 + F4 fused for the finest layer (scale 1): the pyramid image there is just the fixed
// [1 2 1]/4 blur of the frame (REFLECT_101), so blur and polynomial expansion run in one
// kernel straight from the uint8 / f32 frame; T and I never touch global memory.
// Register-blocked: a thread produces 4 vertically adjacent r-values, then 4 horizontally
// adjacent outputs (shared loads as float4), with the tap count a compile-time constant.
// ------------------------------------------------------------------------------------
constexpr int P0_TX = 64, P0_THREADS = 256;
// tile height: as tall as the 48 KB of static shared memory allow (the blur halo shrinks with it)
template <int N>
struct P0Tile {
    static constexpr int TY = N <= 5 ? 24 : 16;
};

template <typename SrcT, int N>
__global__ void __launch_bounds__(P0_THREADS) k_pyr0_polyexp(const SrcT* __restrict__ src, float* __restrict__ R,
                                                             int w, int h, int imgs_per_array, PolyCoef pc) {
    constexpr int P0_TY = P0Tile<N>::TY;
    constexpr int RW = P0_TX + 2 * N, RH = P0_TY + 2 * N;  // blurred-image region
    constexpr int OX = (N + 1 + 3) & ~3;                   // raw region starts OX columns left of the tile (x4)
    constexpr int NWD = (P0_TX + OX + N + 1 + 3) / 4;      // 4-pixel words per raw row
    constexpr int SWS = NWD * 4 + 4, SHS = RH + 2;         // raw frame region (padded stride, 16-byte aligned)
    constexpr int RWP = (RW + 3) & ~3;                     // r-array row stride, float4-aligned
    __shared__ __align__(16) float sS[SHS * SWS];
    __shared__ float sI[RH * RW];
    __shared__ __align__(16) float sr[3][P0_TY * RWP];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * P0_TX, y0 = blockIdx.y * P0_TY, b = blockIdx.z;
    const int ox = x0 - OX, oy = y0 - N - 1;
    const size_t plane = static_cast<size_t>(w) * h;
    const SrcT* sb = src + b * plane;
    // raw frame -> shared.  Tiles whose raw region lies inside the image (all but the border ring)
    // move 4 pixels per load; border tiles reflect (BORDER_REFLECT_101) pixel by pixel.
    if (ox >= 0 && ox + NWD * 4 <= w && oy >= 0 && oy + SHS <= h && (w & 3) == 0) {
        // every thread's loads are issued before the first store (one memory latency per tile, not one per trip)
        constexpr int TRIPS = (SHS * NWD + P0_THREADS - 1) / P0_THREADS;
        float4 v[TRIPS];
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * P0_THREADS;
            if (i < SHS * NWD) {
                const int yy = i / NWD, wd = i - yy * NWD;
                const SrcT* p = sb + static_cast<size_t>(oy + yy) * w + ox + wd * 4;
                if (sizeof(SrcT) == 1) {
                    const uchar4 u = *reinterpret_cast<const uchar4*>(p);
                    v[t] = make_float4(u.x, u.y, u.z, u.w);
                } else {
                    v[t] = *reinterpret_cast<const float4*>(p);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * P0_THREADS;
            if (i < SHS * NWD) {
                const int yy = i / NWD, wd = i - yy * NWD;
                *reinterpret_cast<float4*>(sS + yy * SWS + wd * 4) = v[t];
            }
        }
    } else if ((w & 3) == 0 && (ox & 3) == 0) {
        // border ring: rows reflect as a whole (still one 4-pixel load per word); only the words that
        // straddle the left / right image edge reflect pixel by pixel
        constexpr int TRIPS = (SHS * NWD + P0_THREADS - 1) / P0_THREADS;
        float4 v[TRIPS];
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * P0_THREADS;
            if (i < SHS * NWD) {
                const int yy = i / NWD, wd = i - yy * NWD;
                const int gy = reflect101(oy + yy, h), gx0 = ox + wd * 4;
                const SrcT* row = sb + static_cast<size_t>(gy) * w;
                if (gx0 >= 0 && gx0 + 3 < w) {
                    if (sizeof(SrcT) == 1) {
                        const uchar4 u = *reinterpret_cast<const uchar4*>(row + gx0);
                        v[t] = make_float4(u.x, u.y, u.z, u.w);
                    } else {
                        v[t] = *reinterpret_cast<const float4*>(row + gx0);
                    }
                } else {
                    v[t] = make_float4(load_px(row + reflect101(gx0, w)), load_px(row + reflect101(gx0 + 1, w)),
                                       load_px(row + reflect101(gx0 + 2, w)), load_px(row + reflect101(gx0 + 3, w)));
                }
            }
        }
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * P0_THREADS;
            if (i < SHS * NWD) {
                const int yy = i / NWD, wd = i - yy * NWD;
                *reinterpret_cast<float4*>(sS + yy * SWS + wd * 4) = v[t];
            }
        }
    } else {
        // ragged widths: reflect pixel by pixel, four independent loads per trip
        for (int i0 = tid; i0 < SHS * NWD * 4; i0 += 4 * P0_THREADS) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * P0_THREADS;
                if (i < SHS * NWD * 4) {
                    const int yy = i / (NWD * 4), xx = i - yy * (NWD * 4);
                    const int gy = reflect101(oy + yy, h), gx = reflect101(ox + xx, w);
                    v[u] = load_px(sb + static_cast<size_t>(gy) * w + gx);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * P0_THREADS;
                if (i < SHS * NWD * 4) {
                    const int yy = i / (NWD * 4), xx = i - yy * (NWD * 4);
                    sS[yy * SWS + xx] = v[u];
                }
            }
        }
    }
    __syncthreads();
    // blurred image at replicate-clamped coordinates (what polyExp's border handling reads).
    // Tiles whose blurred region needs no clamping (all but the border ring) walk columns: the row
    // pass of a raw row is computed once and reused by the three blurred rows that read it.
    const bool no_clamp = x0 - N >= 0 && x0 - N + RW <= w && y0 - N >= 0 && y0 - N + RH <= h;
    if (no_clamp) {
        constexpr int CH = 4, RPC = (RH + CH - 1) / CH;  // row chunks per column, rows per chunk
        for (int i = tid; i < CH * RW; i += P0_THREADS) {
            const int ch = i / RW, xx = i - ch * RW;
            const int r0 = ch * RPC;
            const float* c = sS + (r0 + 1) * SWS + (xx + OX - N);   // raw pixel under blurred (r0, xx)
            float tm = 0.25f * c[-SWS - 1] + 0.5f * c[-SWS] + 0.25f * c[-SWS + 1];
            float t0 = 0.25f * c[-1] + 0.5f * c[0] + 0.25f * c[1];
#pragma unroll
            for (int k = 0; k < RPC; ++k) {
                const int yy = r0 + k;
                if (yy >= RH) break;
                c += SWS;
                const float tp = 0.25f * c[-1] + 0.5f * c[0] + 0.25f * c[1];
                sI[yy * RW + xx] = 0.25f * tm + 0.5f * t0 + 0.25f * tp;
                tm = t0;
                t0 = tp;
            }
        }
    } else
    for (int i = tid; i < RH * RW; i += P0_THREADS) {
        const int yy = i / RW, xx = i - yy * RW;
        const int gy = min(max(y0 - N + yy, 0), h - 1), gx = min(max(x0 - N + xx, 0), w - 1);
        const float* c = sS + (gy - oy) * SWS + (gx - ox);
        // rows first (f32), then columns, like the separable filter
        const float t0 = 0.25f * c[-SWS - 1] + 0.5f * c[-SWS] + 0.25f * c[-SWS + 1];
        const float t1 = 0.25f * c[-1] + 0.5f * c[0] + 0.25f * c[1];
        const float t2 = 0.25f * c[SWS - 1] + 0.5f * c[SWS] + 0.25f * c[SWS + 1];
        sI[i] = 0.25f * t0 + 0.5f * t1 + 0.25f * t2;
    }
    __syncthreads();
    // vertical pass: 4 output rows per item
    for (int i = tid; i < (P0_TY / 4) * RW; i += P0_THREADS) {
        const int g = i / RW, xx = i - g * RW;
        float v[4 + 2 * N];
#pragma unroll
        for (int j = 0; j < 4 + 2 * N; ++j) v[j] = sI[(g * 4 + j) * RW + xx];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float r0 = v[o + N] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
            for (int k = 1; k <= N; ++k) {
                const float s0 = v[o + N - k], s1 = v[o + N + k];
                const float p = s0 + s1;
                r0 = fmaf(pc.g[k], p, r0);
                r1 = fmaf(pc.xg[k], s1 - s0, r1);
                r2 = fmaf(pc.xxg[k], p, r2);
            }
            const int a = (g * 4 + o) * RWP + xx;
            sr[0][a] = r0;
            sr[1][a] = r1;
            sr[2][a] = r2;
        }
    }
    __syncthreads();
    // horizontal pass: 4 outputs per thread
    float4* Rq;
    float* Rs;
    r_out(R, imgs_per_array, b, plane, Rq, Rs);
    for (int i = tid; i < P0_TY * (P0_TX / 4); i += P0_THREADS) {
        const int ty = i / (P0_TX / 4), tx = (i - ty * (P0_TX / 4)) * 4;
        const int gy = y0 + ty, gx = x0 + tx;
        if (gy >= h || gx >= w) continue;
        constexpr int NV = ((4 + 2 * N) + 3) & ~3;
        float a0[NV], a1[NV], a2[NV];
#pragma unroll
        for (int j = 0; j < NV; j += 4) {
            *reinterpret_cast<float4*>(a0 + j) = *reinterpret_cast<const float4*>(&sr[0][ty * RWP + tx + j]);
            *reinterpret_cast<float4*>(a1 + j) = *reinterpret_cast<const float4*>(&sr[1][ty * RWP + tx + j]);
            *reinterpret_cast<float4*>(a2 + j) = *reinterpret_cast<const float4*>(&sr[2][ty * RWP + tx + j]);
        }
        float o0[4], o1[4], o2[4], o3[4], o4[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int c = o + N;
            float b1 = a0[c] * pc.g[0], b3 = a1[c] * pc.g[0], b5 = a2[c] * pc.g[0];
            float b2 = 0.f, b4 = 0.f, b6 = 0.f;
#pragma unroll
            for (int k = 1; k <= N; ++k) {
                const float tg = a0[c + k] + a0[c - k];
                b1 = fmaf(tg, pc.g[k], b1);
                b4 = fmaf(tg, pc.xxg[k], b4);
                b2 = fmaf(a0[c + k] - a0[c - k], pc.xg[k], b2);
                b3 = fmaf(a1[c + k] + a1[c - k], pc.g[k], b3);
                b6 = fmaf(a1[c + k] - a1[c - k], pc.xg[k], b6);
                b5 = fmaf(a2[c + k] + a2[c - k], pc.g[k], b5);
            }
            o0[o] = b3 * pc.ig11;
            o1[o] = b2 * pc.ig11;
            o2[o] = fmaf(b1, pc.ig03, b5 * pc.ig33);
            o3[o] = fmaf(b1, pc.ig03, b4 * pc.ig33);
            o4[o] = b6 * pc.ig55;
        }
        const size_t off = static_cast<size_t>(gy) * w + gx;
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (gx + o < w) Rq[off + o] = make_float4(o0[o], o1[o], o2[o], o3[o]);
        if ((w & 3) == 0 && gx + 3 < w) {
            *reinterpret_cast<float4*>(Rs + off) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        } else {
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (gx + o < w) Rs[off + o] = o4[o];
        }
    }
}

// ------------------------------------------------------------------------------------
// The same, with the output staged through shared memory and written by the bulk-copy engine
// (cp.async.bulk, SASS UBLKCP), prev and next frames in one launch.
//
// The kernel is bound by the L1 data stage (shared-memory wavefronts + one wavefront per 128-byte
// line a global store instruction touches).  Two changes against k_pyr0_polyexp:
//   * the horizontal pass gives a thread 8 adjacent outputs (15 16-byte shared loads per 8 pixels
//     instead of 24); a quarter-warp is eight ROWS of one 8-pixel block, and the row strides of the
//     r arrays and of the staging tile are odd numbers of 16-byte chunks, so the 16-byte loads and
//     stores of a quarter-warp fall into eight different bank groups;
//   * a thread's 8 x (float4 + float) results go to a staging tile in shared memory (it takes the
//     place of the raw / blurred image, dead by then) and one lane per row hands the finished row —
//     1 KiB of float4 coefficients, 256 bytes of the fifth — to the bulk-copy engine: the global
//     stores no longer pass through the LSU, where a thread-contiguous 128-byte store costs a
//     wavefront per lane.
// ------------------------------------------------------------------------------------
// Packed f32x2 arithmetic (sm_100: one FFMA2 / FADD2 / FMUL2 issue slot does two IEEE fp32 operations
// on a 64-bit register pair; every lane of the pair is rounded exactly like the scalar instruction).
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(r))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(r))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(r))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(r))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
          "l"(reinterpret_cast<unsigned long long&>(c)));
    return r;
}
__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }

template <int N>
struct P0T {
    static constexpr int TY = P0Tile<N>::TY;
    static constexpr int RW = P0_TX + 2 * N, RH = TY + 2 * N;
    static constexpr int OX = (N + 1 + 3) & ~3;
    static constexpr int NWD = (P0_TX + OX + N + 1 + 3) / 4;
    static constexpr int SWS = NWD * 4 + 4, SHS = RH + 2;
    // r arrays of the horizontal pass: (r0, r1) interleaved per pixel, RW float2 = RW / 2 chunks per row (odd:
    // RW = 74 / 78), and r2 with an odd number of 16-byte chunks per row
    static constexpr int RWP = 4 * ((((RW + 3) / 4)) | 1);
    static constexpr int QS = P0_TX + 1;                     // staging row stride, float4 (65 chunks)
    static constexpr int SS = P0_TX + 4;                     // staging row stride of the fifth coefficient, floats
    static constexpr int A_FLOATS_IN = SHS * SWS + RH * RW;  // raw + blurred image
    static constexpr int A_FLOATS_OUT = TY * QS * 4 + TY * SS;
    static constexpr int A_FLOATS = ((A_FLOATS_IN > A_FLOATS_OUT ? A_FLOATS_IN : A_FLOATS_OUT) + 3) & ~3;
    static constexpr size_t SMEM = static_cast<size_t>(A_FLOATS + TY * (2 * RW + RWP)) * sizeof(float);
    static_assert(TY % 8 == 0 && (TY / 8) * 2 <= P0_THREADS / 32, "horizontal pass: 8 rows x 4 blocks per warp");
    static_assert(RW % 2 == 0 && (RW / 2) % 2 == 1 && (OX - N) % 2 == 1 && SWS % 2 == 0,
                  "column pairs: 8-byte aligned loads, odd chunk strides");
};

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(static_cast<unsigned>(__cvta_generic_to_shared(ssrc))), "r"(bytes)
                 : "memory");
}

template <typename SrcT, int N, int MINB, bool BLUR = true>
__global__ void __launch_bounds__(P0_THREADS, MINB) k_pyr0_polyexp_t(const SrcT* __restrict__ src0,
                                                               const SrcT* __restrict__ src1, int n0,
                                                               float* __restrict__ R, int w, int h,
                                                               int imgs_per_array, PolyCoef pc) {
    using T = P0T<N>;
    constexpr int P0_TY = T::TY, RW = T::RW, RH = T::RH, OX = T::OX, NWD = T::NWD, SWS = T::SWS, SHS = T::SHS,
                  RWP = T::RWP;
    extern __shared__ __align__(128) float smem0[];
    float* sS = smem0;                        // [SHS][SWS] raw frame region
    float* sI = smem0 + SHS * SWS;            // [RH][RW] blurred image
    float2* sR01 = reinterpret_cast<float2*>(smem0 + T::A_FLOATS);   // [P0_TY][RW] (r0, r1) per pixel
    float* sR2 = smem0 + T::A_FLOATS + 2 * P0_TY * RW;                 // [P0_TY][RWP]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * P0_TX, y0 = blockIdx.y * P0_TY, b = blockIdx.z;
    const int ox = x0 - OX, oy = y0 - N - 1;
    const size_t plane = static_cast<size_t>(w) * h;
    // images 0 .. n0-1 come from src0, the rest from src1 (prev and next frames in one launch)
    const SrcT* sb = (b < n0 ? src0 + b * plane : src1 + (b - n0) * plane);
    constexpr int TRIPS = (SHS * NWD + P0_THREADS - 1) / P0_THREADS;
    if (!BLUR) {
        // coarse layers: src already is the pyramid image (f32, any width); replicate-clamped copy
        const float* ib = reinterpret_cast<const float*>(sb);
#pragma unroll 5
        for (int i = tid; i < RH * RW; i += P0_THREADS) {
            const int yy = i / RW, xx = i - yy * RW;
            const int gy = min(max(y0 - N + yy, 0), h - 1), gx = min(max(x0 - N + xx, 0), w - 1);
            sI[i] = __ldg(ib + static_cast<size_t>(gy) * w + gx);
        }
    } else {
    if (ox >= 0 && ox + NWD * 4 <= w && oy >= 0 && oy + SHS <= h) {
        float4 v[TRIPS];
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * P0_THREADS;
            if (i < SHS * NWD) {
                const int yy = i / NWD, wd = i - yy * NWD;
                const SrcT* p = sb + ((oy + yy) * w + ox + wd * 4);   // a frame has fewer than 2^31 pixels
                if (sizeof(SrcT) == 1) {
                    const uchar4 u = *reinterpret_cast<const uchar4*>(p);
                    v[t] = make_float4(u.x, u.y, u.z, u.w);
                } else {
                    v[t] = *reinterpret_cast<const float4*>(p);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * P0_THREADS;
            if (i < SHS * NWD) {
                const int yy = i / NWD, wd = i - yy * NWD;
                *reinterpret_cast<float4*>(sS + yy * SWS + wd * 4) = v[t];
            }
        }
    } else {
        // border ring: rows reflect as a whole (still one 4-pixel load per word); only the words that
        // straddle the left / right image edge reflect pixel by pixel
        float4 v[TRIPS];
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * P0_THREADS;
            if (i < SHS * NWD) {
                const int yy = i / NWD, wd = i - yy * NWD;
                const int gy = reflect101(oy + yy, h), gx0 = ox + wd * 4;
                const SrcT* row = sb + static_cast<size_t>(gy) * w;
                if (gx0 >= 0 && gx0 + 3 < w) {
                    if (sizeof(SrcT) == 1) {
                        const uchar4 u = *reinterpret_cast<const uchar4*>(row + gx0);
                        v[t] = make_float4(u.x, u.y, u.z, u.w);
                    } else {
                        v[t] = *reinterpret_cast<const float4*>(row + gx0);
                    }
                } else {
                    v[t] = make_float4(load_px(row + reflect101(gx0, w)), load_px(row + reflect101(gx0 + 1, w)),
                                       load_px(row + reflect101(gx0 + 2, w)), load_px(row + reflect101(gx0 + 3, w)));
                }
            }
        }
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            const int i = tid + t * P0_THREADS;
            if (i < SHS * NWD) {
                const int yy = i / NWD, wd = i - yy * NWD;
                *reinterpret_cast<float4*>(sS + yy * SWS + wd * 4) = v[t];
            }
        }
    }
    __syncthreads();
    // blurred image at replicate-clamped coordinates (what polyExp's border handling reads)
    const bool no_clamp = x0 - N >= 0 && x0 - N + RW <= w && y0 - N >= 0 && y0 - N + RH <= h;
    if (no_clamp) {
        // two adjacent columns per thread, packed arithmetic; the row pass of a raw row is computed once and
        // reused by the three blurred rows that read it
        constexpr int CH = N <= 5 ? 6 : 4, RPC = (RH + CH - 1) / CH;  // row chunks per column pair, rows per chunk
        const float2 Q = dup2(0.25f), HF = dup2(0.5f);
        // a warp = one row chunk x 32 adjacent column pairs (conflict-free 8-byte accesses); the RW / 2 - 32
        // pairs left over of all chunks share one more warp
        constexpr int REST = RW / 2 - 32;
        static_assert(REST > 0 && CH * REST <= 32 && CH + 1 <= P0_THREADS / 32, "blur: CH warps + one for the rest");
        const int bw = tid >> 5, bl = tid & 31;
        const bool b_on = bw < CH || (bw == CH && bl < CH * REST);
        if (b_on) {
            const int ch = bw < CH ? bw : bl / REST, xx = 2 * (bw < CH ? bl : 32 + bl % REST);
            const int r0 = ch * RPC;
            // raw pixel left of blurred (r0, xx): an even shared-memory column, so column pairs load as float2
            const float* c = sS + (r0 + 1) * SWS + (xx + OX - N - 1);
            auto row_pass = [&](const float* q) {
                const float2 a = *reinterpret_cast<const float2*>(q), b2 = *reinterpret_cast<const float2*>(q + 2);
                return fma2(Q, b2, fma2(HF, make_float2(a.y, b2.x), mul2(Q, a)));
            };
            float2 tm = row_pass(c - SWS), t0 = row_pass(c);
#pragma unroll
            for (int k = 0; k < RPC; ++k) {
                const int yy = r0 + k;
                if (yy >= RH) break;
                c += SWS;
                const float2 tp = row_pass(c);
                *reinterpret_cast<float2*>(sI + yy * RW + xx) = fma2(Q, tp, fma2(HF, t0, mul2(Q, tm)));
                tm = t0;
                t0 = tp;
            }
        }
    } else {
        for (int i = tid; i < RH * RW; i += P0_THREADS) {
            const int yy = i / RW, xx = i - yy * RW;
            const int gy = min(max(y0 - N + yy, 0), h - 1), gx = min(max(x0 - N + xx, 0), w - 1);
            const float* c = sS + (gy - oy) * SWS + (gx - ox);
            // the operation order of the packed path above
            const float t0 = fmaf(0.25f, c[-SWS + 1], fmaf(0.5f, c[-SWS], 0.25f * c[-SWS - 1]));
            const float t1 = fmaf(0.25f, c[1], fmaf(0.5f, c[0], 0.25f * c[-1]));
            const float t2 = fmaf(0.25f, c[SWS + 1], fmaf(0.5f, c[SWS], 0.25f * c[SWS - 1]));
            sI[i] = fmaf(0.25f, t2, fmaf(0.5f, t1, 0.25f * t0));
        }
    }
    }   // BLUR
    __syncthreads();
    // vertical pass: 4 output rows x 2 adjacent columns per item, packed arithmetic; the same warp mapping
    // (a warp = one group of 4 rows x 32 adjacent column pairs, the rest of all groups in one more warp)
    constexpr int VG = P0_TY / 4, VREST = RW / 2 - 32;
    static_assert(VG * VREST <= 32 && VG + 1 <= P0_THREADS / 32, "vertical pass: VG warps + one for the rest");
    const int vw = tid >> 5, vl = tid & 31;
    if (vw < VG || (vw == VG && vl < VG * VREST)) {
        const int g = vw < VG ? vw : vl / VREST, xx = 2 * (vw < VG ? vl : 32 + vl % VREST);
        float2 v[4 + 2 * N];
#pragma unroll
        for (int j = 0; j < 4 + 2 * N; ++j) v[j] = *reinterpret_cast<const float2*>(sI + (g * 4 + j) * RW + xx);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float2 r0 = mul2(v[o + N], dup2(pc.g[0])), r1 = make_float2(0.f, 0.f), r2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 1; k <= N; ++k) {
                const float2 s0 = v[o + N - k], s1 = v[o + N + k];
                const float2 p = add2(s0, s1);
                r0 = fma2(dup2(pc.g[k]), p, r0);
                r1 = fma2(dup2(pc.xg[k]), sub2(s1, s0), r1);
                r2 = fma2(dup2(pc.xxg[k]), p, r2);
            }
            const int row = g * 4 + o;
            *reinterpret_cast<float4*>(sR01 + row * RW + xx) = make_float4(r0.x, r1.x, r0.y, r1.y);
            *reinterpret_cast<float2*>(sR2 + row * RWP + xx) = r2;
        }
    }
    __syncthreads();   // the raw / blurred image is dead from here on: its place becomes the staging tile
    float4* stQ = reinterpret_cast<float4*>(smem0);          // [P0_TY][QS]
    float* stS = smem0 + P0_TY * T::QS * 4;                  // [P0_TY][SS]
    // horizontal pass: 8 outputs per thread; a quarter-warp = 8 rows of one 8-pixel block
    {
        const int wi = tid >> 5, lane = tid & 31;
        const int ty = (wi >> 1) * 8 + (lane & 7), blk = (wi & 1) * 4 + (lane >> 3);
        if (ty < P0_TY) {
            const int tx = blk * 8;
            constexpr int NP = 8 + 2 * N;              // pixels of the window of 8 outputs
            constexpr int NV = (NP + 3) & ~3;
            float2 a01[NP];                            // (r0, r1) per pixel
            float a2[NV];
#pragma unroll
            for (int j = 0; j < NP; j += 2) {
                const float4 q = *reinterpret_cast<const float4*>(sR01 + ty * RW + tx + j);
                a01[j] = make_float2(q.x, q.y), a01[j + 1] = make_float2(q.z, q.w);
            }
#pragma unroll
            for (int j = 0; j < NV; j += 4)
                *reinterpret_cast<float4*>(a2 + j) = *reinterpret_cast<const float4*>(sR2 + ty * RWP + tx + j);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float o4[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int c = half * 4 + o + N;
                    float2 b13 = mul2(a01[c], dup2(pc.g[0]));          // (b1, b3)
                    float2 b26 = make_float2(0.f, 0.f);                // (b2, b6)
                    float b5 = a2[c] * pc.g[0], b4 = 0.f;
#pragma unroll
                    for (int k = 1; k <= N; ++k) {
                        const float2 tg = add2(a01[c + k], a01[c - k]);
                        b13 = fma2(tg, dup2(pc.g[k]), b13);
                        b26 = fma2(sub2(a01[c + k], a01[c - k]), dup2(pc.xg[k]), b26);
                        b4 = fmaf(tg.x, pc.xxg[k], b4);
                        b5 = fmaf(a2[c + k] + a2[c - k], pc.g[k], b5);
                    }
                    const float b1 = b13.x, b3 = b13.y, b2 = b26.x, b6 = b26.y;
                    stQ[ty * T::QS + tx + half * 4 + o] =
                        make_float4(b3 * pc.ig11, b2 * pc.ig11, fmaf(b1, pc.ig03, b5 * pc.ig33), fmaf(b1, pc.ig03, b4 * pc.ig33));
                    o4[o] = b6 * pc.ig55;
                }
                *reinterpret_cast<float4*>(stS + ty * T::SS + tx + half * 4) = make_float4(o4[0], o4[1], o4[2], o4[3]);
            }
        }
    }
    // generic-proxy writes to shared memory become visible to the bulk-copy engine
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int npx = min(P0_TX, w - x0);
    // rows of the fifth coefficient start 16-byte aligned only when w % 4 == 0 (always, with BLUR: the launcher
    // checks); otherwise they leave through the LSU
    const bool s_bulk = BLUR || (w & 3) == 0;
    // lane l of warp v hands row v + 8 l to the copy engine (three lanes per warp issue side by side; one lane
    // looping over the warp's rows cost 6 % of the kernel's issue slots)
    const int ty_b = (tid >> 5) + (P0_THREADS / 32) * (tid & 31);
    const bool mover = ty_b < P0_TY && y0 + ty_b < h;
    if (!mover && s_bulk) return;
    float4* Rq;
    float* Rs;
    r_out(R, imgs_per_array, b, plane, Rq, Rs);
    if (mover) {
        const int off = (y0 + ty_b) * w + x0;
        bulk_store(Rq + off, stQ + ty_b * T::QS, npx * 16);
        if (s_bulk) bulk_store(Rs + off, stS + ty_b * T::SS, npx * 4);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (!s_bulk) {
        for (int i = tid; i < P0_TY * P0_TX; i += P0_THREADS) {
            const int ty = i / P0_TX, tx = i - ty * P0_TX;
            if (y0 + ty < h && tx < npx) Rs[(y0 + ty) * w + x0 + tx] = stS[ty * T::SS + tx];
        }
    }
    // shared memory must stay intact until the engine has read it
    if (mover) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------
// F3: flow initialisation of a finer layer = bilinear resize of the coarser flow * mul
// ------------------------------------------------------------------------------------
// Source taps and weights come from per-layer tables (up_tables): the fp64 tap arithmetic that
// decides which texels are read runs once per column / row on the host, not per pixel.
// A thread produces UP_ROWS rows x 2 columns (one 16-byte store per row): few, fat CTAs — a CTA per
// 128 single-pixel threads spent its life waiting for one dependent load.
constexpr int UP_ROWS = 4;
__global__ void __launch_bounds__(256) k_upsample_flow(const float2* __restrict__ fin, float2* __restrict__ fout,
                                                       int hi, int wi, int ho, int wo, const int* __restrict__ sxt,
                                                       const float* __restrict__ fxt, const int* __restrict__ syt,
                                                       const float* __restrict__ fyt, float fmul) {
    const int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y0 = blockIdx.y * UP_ROWS, b = blockIdx.z;
    if (x >= wo) return;
    const bool two = x + 1 < wo;
    const int xb = two ? x + 1 : x;
    const int sxa = sxt[x], sxb = sxt[xb];
    const int sxa1 = min(sxa + 1, wi - 1), sxb1 = min(sxb + 1, wi - 1);
    const float ga1 = fxt[x], ga0 = 1.f - ga1, gb1 = fxt[xb], gb0 = 1.f - gb1;
    const float2* p = fin + static_cast<size_t>(b) * hi * wi;
    float2* o = fout + (static_cast<size_t>(b) * ho + y0) * wo + x;
    const bool vec = two && (wo & 1) == 0;  // 16-byte aligned pair
#pragma unroll
    for (int rr = 0; rr < UP_ROWS; ++rr, o += wo) {
        const int y = y0 + rr;
        if (y >= ho) break;
        const int sy = syt[y], sy1 = min(sy + 1, hi - 1);
        const float gy1 = fyt[y], gy0 = 1.f - gy1;
        const float2* r0 = p + sy * wi;
        const float2* r1 = p + sy1 * wi;
        // the interpolation itself in f32 — horizontal pass rounded to f32, then vertical, as cv::resize does
        const float2 a = r0[sxa], a1 = r0[sxa1], c = r1[sxa], c1 = r1[sxa1];
        const float2 e = r0[sxb], e1 = r0[sxb1], g = r1[sxb], g1 = r1[sxb1];
        float2 oa, ob;
        oa.x = (gy0 * (ga0 * a.x + ga1 * a1.x) + gy1 * (ga0 * c.x + ga1 * c1.x)) * fmul;
        oa.y = (gy0 * (ga0 * a.y + ga1 * a1.y) + gy1 * (ga0 * c.y + ga1 * c1.y)) * fmul;
        ob.x = (gy0 * (gb0 * e.x + gb1 * e1.x) + gy1 * (gb0 * g.x + gb1 * g1.x)) * fmul;
        ob.y = (gy0 * (gb0 * e.y + gb1 * e1.y) + gy1 * (gb0 * g.y + gb1 * g1.y)) * fmul;
        if (vec) {
            *reinterpret_cast<float4*>(o) = make_float4(oa.x, oa.y, ob.x, ob.y);
        } else {
            o[0] = oa;
            if (two) o[1] = ob;
        }
    }
}

// ------------------------------------------------------------------------------------
// F5: updateMatrices for one pixel
// ------------------------------------------------------------------------------------
__constant__ float c_border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};

// Split in two so a caller can issue the loads of several pixels before consuming any
// (the gathers are the long-latency part of a flow iteration).
struct MTaps {
    float q[5];      // R0 at the pixel
    float t[5][4];   // R1 at the four bilinear taps, per channel
    float a00, a01, a10, a11;
    float dx, dy;
    bool inside;
};

__device__ __forceinline__ void m_gather(const RView& r0, const RView& r1, int w, int h, int x, int y, float2 f,
                                         MTaps& T) {
    T.dx = f.x;
    T.dy = f.y;
    float fx = static_cast<float>(x) + f.x, fy = static_cast<float>(y) + f.y;
    const float x1f = floorf(fx), y1f = floorf(fy);
    fx -= x1f;
    fy -= y1f;
    const int o = y * w + x;
    const float4 q = r0.q[o];
    T.q[0] = q.x, T.q[1] = q.y, T.q[2] = q.z, T.q[3] = q.w;
    T.q[4] = r0.s[o];
    // float-domain test also rejects NaN / huge displacements
    T.inside = x1f >= 0.f && x1f < static_cast<float>(w - 1) && y1f >= 0.f && y1f < static_cast<float>(h - 1);
    if (T.inside) {
        const int i00 = static_cast<int>(y1f) * w + static_cast<int>(x1f);
        T.a01 = fx * (1.f - fy);
        T.a11 = fx * fy;
        T.a00 = (1.f - fx) * (1.f - fy);
        T.a10 = (1.f - fx) * fy;
        const float4* pq = r1.q + i00;
        const float* ps = r1.s + i00;
        const float4 t0 = pq[0], t1 = pq[1], t2 = pq[w], t3 = pq[w + 1];
        T.t[0][0] = t0.x, T.t[1][0] = t0.y, T.t[2][0] = t0.z, T.t[3][0] = t0.w;
        T.t[0][1] = t1.x, T.t[1][1] = t1.y, T.t[2][1] = t1.z, T.t[3][1] = t1.w;
        T.t[0][2] = t2.x, T.t[1][2] = t2.y, T.t[2][2] = t2.z, T.t[3][2] = t2.w;
        T.t[0][3] = t3.x, T.t[1][3] = t3.y, T.t[2][3] = t3.z, T.t[3][3] = t3.w;
        T.t[4][0] = ps[0];
        T.t[4][1] = ps[1];
        T.t[4][2] = ps[w];
        T.t[4][3] = ps[w + 1];
    }
}

template <bool BORDER = true>
__device__ __forceinline__ void m_finish(const MTaps& T, int w, int h, int x, int y, float M[5]) {
    const float dx = T.dx, dy = T.dy;
    float r2, r3, r4, r5, r6;
    if (T.inside) {
        float s[5];
#pragma unroll
        for (int c = 0; c < 5; ++c)
            s[c] = T.a00 * T.t[c][0] + T.a01 * T.t[c][1] + T.a10 * T.t[c][2] + T.a11 * T.t[c][3];
        r2 = s[0];
        r3 = s[1];
        r4 = (T.q[2] + s[2]) * 0.5f;
        r5 = (T.q[3] + s[3]) * 0.5f;
        r6 = (T.q[4] + s[4]) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = T.q[2];
        r5 = T.q[3];
        r6 = T.q[4] * 0.5f;
    }
    r2 = (T.q[0] - r2) * 0.5f;
    r3 = (T.q[1] - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if (BORDER && (x < 5 || x >= w - 5 || y < 5 || y >= h - 5)) {
        float sc = (x < 5 ? c_border[x] : 1.f) * (x >= w - 5 ? c_border[w - 1 - x] : 1.f) *
                   (y < 5 ? c_border[y] : 1.f) * (y >= h - 5 ? c_border[h - 1 - y] : 1.f);
        r2 *= sc, r3 *= sc, r4 *= sc, r5 *= sc, r6 *= sc;
    }
    M[0] = r4 * r4 + r6 * r6;
    M[1] = (r4 + r5) * r6;
    M[2] = r5 * r5 + r6 * r6;
    M[3] = r4 * r2 + r6 * r3;
    M[4] = r6 * r2 + r5 * r3;
}

__device__ __forceinline__ void compute_M(const RView& r0, const RView& r1, int w, int h, int x, int y, float2 f,
                                          float M[5]) {
    MTaps T;
    m_gather(r0, r1, w, h, x, y, f, T);
    m_finish(T, w, h, x, y, M);
}

__global__ void __launch_bounds__(256) k_update_matrices(const float* __restrict__ R0, const float* __restrict__ R1,
                                                         const float2* __restrict__ flow, float* __restrict__ Mout,
                                                         int w, int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y, b = blockIdx.z;
    if (x >= w) return;
    const size_t plane = static_cast<size_t>(w) * h;
    float M[5];
    compute_M(r_view(R0, gridDim.z, b, plane), r_view(R1, gridDim.z, b, plane), w, h, x, y,
              flow[b * plane + static_cast<size_t>(y) * w + x], M);
    float* o = Mout + b * 5 * plane + static_cast<size_t>(y) * w + x;
#pragma unroll
    for (int c = 0; c < 5; ++c) o[c * plane] = M[c];
}

// ------------------------------------------------------------------------------------
// F5 + F6 fused: one flow iteration.  A CTA owns a FI_TX x FI_TY tile of output pixels,
// builds M for the tile plus an m-pixel halo in shared memory (FUSED: straight from
// R0 / R1 / flow; otherwise from a precomputed M field), box-sums it separably with
// direct (non-running) sums and solves the 2x2 system per pixel.
// ------------------------------------------------------------------------------------
constexpr int FI_TX = 32, FI_TY = 32, FI_THREADS = 256;
#ifndef DATMO_FI_TILE_DEFAULT
#define DATMO_FI_TILE_DEFAULT 1
#endif

__device__ __forceinline__ float2 solve_flow(const float g[5]) {
    // g = blurred (g11, g12, g22, h1, h2)
    float idet = 1.f / (g[0] * g[2] - g[1] * g[1] + 1e-3f);
    float2 o;
    o.x = (g[0] * g[4] - g[1] * g[3]) * idet;
    o.y = (g[2] * g[3] - g[1] * g[4]) * idet;
    return o;
}

template <bool FUSED>
__global__ void __launch_bounds__(FI_THREADS) k_flow_iter(const float* __restrict__ R0, const float* __restrict__ R1,
                                                          const float2* __restrict__ flow_in,
                                                          const float* __restrict__ Min, float2* __restrict__ flow_out,
                                                          int w, int h, int m, float norm) {
    extern __shared__ float smem[];
    const int RW = FI_TX + 2 * m, RH = FI_TY + 2 * m;
    const int SW = RW | 1;  // odd row stride
    float* sM = smem;                       // [5][RH][SW]
    float* sV = sM + 5 * RH * SW;           // [5][FI_TY][SW]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * FI_TX, y0 = blockIdx.y * FI_TY, b = blockIdx.z;
    const size_t plane = static_cast<size_t>(w) * h;
    const RView R0b = r_view(R0, gridDim.z, b, plane), R1b = r_view(R1, gridDim.z, b, plane);
    const float* Mb = Min + static_cast<size_t>(b) * 5 * plane;
    const float2* fb = flow_in + static_cast<size_t>(b) * plane;
    // phase 1: M over the tile + halo (replicate border = clamp the coordinates)
    for (int i = tid; i < RH * RW; i += FI_THREADS) {
        int yy = i / RW, xx = i - yy * RW;
        int gx = min(max(x0 - m + xx, 0), w - 1);
        int gy = min(max(y0 - m + yy, 0), h - 1);
        float M[5];
        if (FUSED) {
            compute_M(R0b, R1b, w, h, gx, gy, fb[static_cast<size_t>(gy) * w + gx], M);
        } else {
#pragma unroll
            for (int c = 0; c < 5; ++c) M[c] = Mb[c * plane + static_cast<size_t>(gy) * w + gx];
        }
#pragma unroll
        for (int c = 0; c < 5; ++c) sM[(c * RH + yy) * SW + xx] = M[c];
    }
    __syncthreads();
    // phase 2: vertical window sums
    const int win = 2 * m + 1;
    for (int i = tid; i < 5 * FI_TY * RW; i += FI_THREADS) {
        int xx = i % RW;
        int t = i / RW;
        int y = t % FI_TY, c = t / FI_TY;
        const float* col = sM + (c * RH + y) * SW + xx;
        float s = 0.f;
        for (int d = 0; d < win; ++d) s += col[d * SW];
        sV[(c * FI_TY + y) * SW + xx] = s;
    }
    __syncthreads();
    // phase 3: horizontal window sums + solve
    float2* fo = flow_out + static_cast<size_t>(b) * plane;
    for (int i = tid; i < FI_TX * FI_TY; i += FI_THREADS) {
        int y = i / FI_TX, x = i - y * FI_TX;
        int gx = x0 + x, gy = y0 + y;
        if (gx >= w || gy >= h) continue;
        float g[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const float* row = sV + (c * FI_TY + y) * SW + x;
            float s = 0.f;
            for (int d = 0; d < win; ++d) s += row[d];
            g[c] = s * norm;
        }
        fo[static_cast<size_t>(gy) * w + gx] = solve_flow(g);
    }
}

// ------------------------------------------------------------------------------------
// F5 + F6 fused, specialised for a compile-time window (winsize 14 / 15 -> 15 taps).
//
// One shared-memory buffer, three in-place passes:
//   1. M for the tile + halo (FUSED: gathered straight from R0 / R1 / flow);
//   2. vertical window sums, one thread per (column, channel), van Herk / Gil-Werman:
//      the column is cut into segments of WIN rows; a window that starts inside segment k
//      is (suffix sum of segment k) + (prefix sum of segment k+1).  Every output is a sum
//      of exactly its own window's values — no running-sum cancellation — at ~3 adds and
//      ~1.5 shared loads per output instead of WIN of each.  The thread owns its column,
//      so the sums overwrite the M values it has already consumed;
//   3. horizontal window sums, one thread per (row, channel), same scheme along x;
//   4. per-pixel 2x2 solve, float2 store.
// ------------------------------------------------------------------------------------
template <int WIN, int NOUT>
__device__ __forceinline__ void window_sums_inplace(float* __restrict__ p, const int stride) {
    constexpr int NIN = NOUT + WIN - 1;
    float S[WIN];
#pragma unroll
    for (int j = 0; j < WIN; ++j) S[j] = p[j * stride];
#pragma unroll
    for (int j = WIN - 2; j >= 0; --j) S[j] += S[j + 1];
#pragma unroll
    for (int base = 0; base < NOUT; base += WIN) {
        p[base * stride] = S[0];
        float P = 0.f;
        float nxt[WIN];
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const int idx = base + WIN + j;
            if (idx < NIN) {
                const float v = p[idx * stride];
                nxt[j] = v;
                P += v;
            } else {
                nxt[j] = 0.f;
            }
            const int y = base + 1 + j;
            if (j < WIN - 1 && y < NOUT) p[y * stride] = S[j + 1] + P;
        }
#pragma unroll
        for (int j = WIN - 2; j >= 0; --j) nxt[j] += nxt[j + 1];
#pragma unroll
        for (int j = 0; j < WIN; ++j) S[j] = nxt[j];
    }
}

template <int TX, int TY, int WIN, int NT_>
struct FlowTile {
    static constexpr int M = WIN / 2;
    static constexpr int RW = TX + 2 * M, RH = TY + 2 * M;
    static constexpr int SW = RW | 1;  // odd row stride: column walks and row walks are both conflict-free
    static constexpr int THREADS = NT_;
    static constexpr size_t SMEM = static_cast<size_t>(5) * RH * SW * sizeof(float);
};

template <int TX, int TY, int WIN, int NT, int MINB, bool FUSED>
__global__ void __launch_bounds__(NT, MINB) k_flow_iter_w(const float* __restrict__ R0, const float* __restrict__ R1,
                                                    const float2* __restrict__ flow_in,
                                                    const float* __restrict__ Min, float2* __restrict__ flow_out,
                                                    int w, int h, float norm) {
    using T = FlowTile<TX, TY, WIN, NT>;
    constexpr int M = T::M, RW = T::RW, RH = T::RH, SW = T::SW;
    extern __shared__ float sM[];  // [5][RH][SW]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, b = blockIdx.z;
    const size_t plane = static_cast<size_t>(w) * h;
    const RView R0b = r_view(R0, gridDim.z, b, plane), R1b = r_view(R1, gridDim.z, b, plane);
    const float* Mb = Min + static_cast<size_t>(b) * 5 * plane;
    const float2* fb = flow_in + static_cast<size_t>(b) * plane;
    // pass 1: M over tile + halo, two pixels per trip: the flow of the NEXT trip is
    // prefetched, and both pixels' gathers are issued before either is consumed.
    {
        constexpr int NPIX = RW * RH;
        auto locate = [&](int i, int& gx, int& gy, int& so) {
            const int yy = i / RW, xx = i - yy * RW;  // RW is a compile-time constant
            gx = min(max(x0 - M + xx, 0), w - 1);
            gy = min(max(y0 - M + yy, 0), h - 1);
            so = yy * SW + xx;
        };
        if (FUSED) {
            // 2-D mapping: LW lanes along x, the thread walks the region in (LW, LH) strides, so
            // coordinates are adds; tiles that touch neither the image border nor its 5-px
            // attenuation band (the common case) skip every clamp and the border test.
            constexpr int LW = 16, LH = NT / LW;
            constexpr int NI = (RW + LW - 1) / LW, NJ = (RH + LH - 1) / LH;
            const int tx = tid % LW, ty = tid / LW;
            const bool interior = x0 - M >= 5 && x0 - M + RW <= w - 5 && y0 - M >= 5 && y0 - M + RH <= h - 5;
            auto body = [&](auto interior_tag) {
                constexpr bool INT = decltype(interior_tag)::value;
                for (int j = 0; j < NJ; ++j) {
                    const int yy = ty + j * LH;
                    if (yy >= RH) break;
                    const int gy = INT ? y0 - M + yy : min(max(y0 - M + yy, 0), h - 1);
                    const float2* frow = fb + static_cast<size_t>(gy) * w;
                    float2 f[NI];
                    int gxs[NI];
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        const int xx = tx + i * LW;
                        gxs[i] = INT ? x0 - M + xx : min(max(x0 - M + xx, 0), w - 1);
                        if (xx < RW) f[i] = frow[gxs[i]];
                    }
                    float* srow = sM + yy * SW + tx;
#pragma unroll
                    for (int i = 0; i < NI; i += 2) {
                        const bool has_a = tx + i * LW < RW;
                        const bool has_b = i + 1 < NI && tx + (i + 1) * LW < RW;
                        MTaps Ta, Tb;
                        if (has_a) m_gather(R0b, R1b, w, h, gxs[i], gy, f[i], Ta);
                        if (has_b) m_gather(R0b, R1b, w, h, gxs[i + 1 < NI ? i + 1 : i], gy, f[i + 1 < NI ? i + 1 : i], Tb);
                        float Mv[5];
                        if (has_a) {
                            m_finish<!INT>(Ta, w, h, gxs[i], gy, Mv);
#pragma unroll
                            for (int c = 0; c < 5; ++c) srow[c * RH * SW + i * LW] = Mv[c];
                        }
                        if (has_b) {
                            m_finish<!INT>(Tb, w, h, gxs[i + 1 < NI ? i + 1 : i], gy, Mv);
#pragma unroll
                            for (int c = 0; c < 5; ++c) srow[c * RH * SW + (i + 1) * LW] = Mv[c];
                        }
                    }
                }
            };
            if (interior)
                body(std::true_type{});
            else
                body(std::false_type{});
        } else {
            for (int i = tid; i < NPIX; i += NT) {
                int gx, gy, so;
                locate(i, gx, gy, so);
#pragma unroll
                for (int c = 0; c < 5; ++c) sM[c * RH * SW + so] = Mb[c * plane + static_cast<size_t>(gy) * w + gx];
            }
        }
    }
    __syncthreads();
    // pass 2: vertical, thread per (channel, column)
    for (int i = tid; i < 5 * RW; i += NT) {
        const int c = i / RW, xx = i - c * RW;
        window_sums_inplace<WIN, TY>(sM + c * RH * SW + xx, SW);
    }
    __syncthreads();
    // pass 3: horizontal, thread per (channel, row); rows 0..TY-1 now hold the vertical sums
    for (int i = tid; i < 5 * TY; i += NT) {
        const int c = i / TY, y = i - c * TY;
        window_sums_inplace<WIN, TX>(sM + (c * RH + y) * SW, 1);
    }
    __syncthreads();
    // pass 4: solve
    float2* fo = flow_out + static_cast<size_t>(b) * plane;
    for (int i = tid; i < TX * TY; i += NT) {
        const int y = i / TX, x = i - y * TX;
        const int gx = x0 + x, gy = y0 + y;
        if (gx >= w || gy >= h) continue;
        float g[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) g[c] = sM[(c * RH + y) * SW + x] * norm;
        fo[static_cast<size_t>(gy) * w + gx] = solve_flow(g);
    }
}

// ------------------------------------------------------------------------------------
// F5 + F6 fused, x-marching form (winsize 14 / 15).
//
// A CTA owns a band of TY output rows and walks a segment of it left to right in groups of
// 32 columns.  Per group:
//   M   M for the 32 NEW columns x (TY + 14) rows (a warp = 32 consecutive pixels of a row, so
//       the R0 / flow loads are whole aligned lines); only the rows are halo — no column is
//       ever evaluated twice inside a segment;
//   V   vertical 15-sums of the new columns (van Herk, thread per (channel, column, half band))
//       appended to a 46-column window of vertical sums whose first 14 columns are carried
//       over from the previous groups;
//   H   horizontal 15-sums over the window -> 32 finished output columns (they trail the new
//       group by 7), in place: output j lands on window column 14 + j and the row's last 14
//       window columns move to the front (the thread that owns a (channel, row) does both);
//   S   2x2 solve straight from the window, float2 store; it shares a barrier interval with the
//       next group's M phase, so a group costs three barriers.
// M evaluations per output pixel: (TY + 14) / TY x (seg + 16) / seg  (1.35 for TY 46, seg 512)
// against 1.75 for the 64 x 32 tile kernel above, and the vertical pass never touches a halo
// column.  Every window sum is still the sum of exactly its own 15 values.  The kernel is bound
// by the L1 data stage (gathers + shared-memory window sums) and by issue, not by HBM: DESIGN.md §5/§6.
// ------------------------------------------------------------------------------------
template <int WIN, int NOUT>
__device__ __forceinline__ void window_sums(const float* __restrict__ in, const int si, float* __restrict__ out,
                                            const int so) {
    // inputs are read one block of WIN ahead of their use, WIN independent loads at a time, so the
    // add chains never wait on shared-memory latency
    constexpr int NIN = NOUT + WIN - 1;
    float S[WIN], nxt[WIN];
#pragma unroll
    for (int j = 0; j < WIN; ++j) S[j] = in[j * si];
#pragma unroll
    for (int j = 0; j < WIN; ++j) nxt[j] = WIN + j < NIN ? in[(WIN + j) * si] : 0.f;
#pragma unroll
    for (int j = WIN - 2; j >= 0; --j) S[j] += S[j + 1];
#pragma unroll
    for (int base = 0; base < NOUT; base += WIN) {
        float ahead[WIN];
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const int idx = base + 2 * WIN + j;
            ahead[j] = (base + WIN < NOUT && idx < NIN) ? in[idx * si] : 0.f;
        }
        out[base * so] = S[0];
        float P = 0.f;
#pragma unroll
        for (int j = 0; j < WIN - 1; ++j) {
            P += nxt[j];
            const int y = base + 1 + j;
            if (y < NOUT) out[y * so] = S[j + 1] + P;
        }
#pragma unroll
        for (int j = WIN - 2; j >= 0; --j) nxt[j] += nxt[j + 1];
#pragma unroll
        for (int j = 0; j < WIN; ++j) S[j] = nxt[j], nxt[j] = ahead[j];
    }
}

// In-place form for the horizontal pass over a row v[0 .. NOUT + WIN - 2] (stride 1): output j
// lands on v[WIN - 1 + j] and the last WIN - 1 inputs move to the front, i.e. the row is ready to
// receive the next NOUT new values behind them.  Safe because a block's outputs are stored only
// after the inputs two blocks ahead are in registers, and everything is program-ordered.
template <int WIN, int NOUT>
__device__ __forceinline__ void window_sums_carry(float* v) {
    constexpr int NIN = NOUT + WIN - 1;
    float S[WIN], nxt[WIN], carry[WIN - 1];
#pragma unroll
    for (int j = 0; j < WIN; ++j) S[j] = v[j];
#pragma unroll
    for (int j = 0; j < WIN; ++j) nxt[j] = WIN + j < NIN ? v[WIN + j] : 0.f;
#pragma unroll
    for (int j = 0; j < WIN - 1; ++j) carry[j] = v[NOUT + j];
#pragma unroll
    for (int j = WIN - 2; j >= 0; --j) S[j] += S[j + 1];
#pragma unroll
    for (int base = 0; base < NOUT; base += WIN) {
        float ahead[WIN];
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const int idx = base + 2 * WIN + j;
            ahead[j] = (base + WIN < NOUT && idx < NIN) ? v[idx] : 0.f;
        }
        v[WIN - 1 + base] = S[0];
        float P = 0.f;
#pragma unroll
        for (int j = 0; j < WIN - 1; ++j) {
            P += nxt[j];
            const int y = base + 1 + j;
            if (y < NOUT) v[WIN - 1 + y] = S[j + 1] + P;
        }
#pragma unroll
        for (int j = WIN - 2; j >= 0; --j) nxt[j] += nxt[j + 1];
#pragma unroll
        for (int j = 0; j < WIN; ++j) S[j] = nxt[j], nxt[j] = ahead[j];
    }
#pragma unroll
    for (int j = 0; j < WIN - 1; ++j) v[j] = carry[j];
}

template <int TY_, int NT_, int VSPLIT_, int MINB_, int PIX_>
struct XmTile {
    static constexpr int TY = TY_, NT = NT_, VSPLIT = VSPLIT_, MINB = MINB_, PIX = PIX_;
    static constexpr int HM = 7, WIN = 15;
    static constexpr int RH = TY + 2 * HM;       // M rows of a group
    static constexpr int MS = 33;                // row stride of the M / G buffer (32 columns + 1)
    static constexpr int VS = 47;                // vertical-sum window: 14 carried + 32 new columns = 46; odd stride
    static constexpr int M_FLOATS = 5 * RH * MS, V_FLOATS = 5 * TY * VS;
    static constexpr size_t SMEM = static_cast<size_t>(M_FLOATS + V_FLOATS) * sizeof(float);
    static_assert(TY % VSPLIT == 0, "band must split evenly for the vertical pass");
};

// base + index as ONE instruction (IMAD.WIDE); left to itself the compiler sign-extends the
// index and builds the address with a four-instruction carry chain.
template <int BYTES, typename P>
__device__ __forceinline__ const P* at_index(const P* base, int i) {
    unsigned long long r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(i), "n"(BYTES), "l"(base));
    return reinterpret_cast<const P*>(r);
}

// Everything one M evaluation gathers.  The R1 taps are loaded unconditionally (index 0 when the
// displaced position leaves the image) and replaced by selects afterwards: no divergent branch,
// so the loads of all pixels of a trip are issued back to back.
struct XmTaps {
    float4 q, t0, t1, t2, t3;
    float s, u0, u1, u2, u3;
    float fx, fy, dx, dy;
    bool inside;
};

__device__ __forceinline__ void xm_gather(const float4* r0q, const float* r0s, const float4* r1q, const float* r1s,
                                          int w, int h, float xf, int gx, int gy, float2 f, XmTaps& T) {
    const int o = gy * w + gx;
    T.q = __ldg(at_index<16>(r0q, o));
    T.s = __ldg(at_index<4>(r0s, o));
    T.dx = f.x, T.dy = f.y;
    const float px = xf + f.x, py = static_cast<float>(gy) + f.y;
    const int ix = __float2int_rd(px), iy = __float2int_rd(py);   // saturating; NaN -> 0
    T.inside = static_cast<unsigned>(ix) < static_cast<unsigned>(w - 1) &&
               static_cast<unsigned>(iy) < static_cast<unsigned>(h - 1);
    T.fx = px - static_cast<float>(ix);
    T.fy = py - static_cast<float>(iy);
    const int i00 = T.inside ? iy * w + ix : 0;
    const float4* pq = at_index<16>(r1q, i00);
    const float* ps = at_index<4>(r1s, i00);
    const float4* pq2 = at_index<16>(r1q, i00 + w);
    const float* ps2 = at_index<4>(r1s, i00 + w);
    T.t0 = __ldg(pq), T.t1 = __ldg(pq + 1), T.t2 = __ldg(pq2), T.t3 = __ldg(pq2 + 1);
    T.u0 = __ldg(ps), T.u1 = __ldg(ps + 1), T.u2 = __ldg(ps2), T.u3 = __ldg(ps2 + 1);
}

// SURVEY.md §3.2 F5 with the out-of-image case folded into selects: r4 = (q2 + q2) / 2 = q2 etc. are exact
__device__ __forceinline__ void xm_finish(const XmTaps& T, bool border, int w, int h, int gx, int gy, float M[5]) {
    const float gx1 = 1.f - T.fx, gy1 = 1.f - T.fy;
    const float a00 = gx1 * gy1, a01 = T.fx * gy1, a10 = gx1 * T.fy, a11 = T.fx * T.fy;
    float s0 = a00 * T.t0.x + a01 * T.t1.x + a10 * T.t2.x + a11 * T.t3.x;
    float s1 = a00 * T.t0.y + a01 * T.t1.y + a10 * T.t2.y + a11 * T.t3.y;
    float s2 = a00 * T.t0.z + a01 * T.t1.z + a10 * T.t2.z + a11 * T.t3.z;
    float s3 = a00 * T.t0.w + a01 * T.t1.w + a10 * T.t2.w + a11 * T.t3.w;
    float s4 = a00 * T.u0 + a01 * T.u1 + a10 * T.u2 + a11 * T.u3;
    if (!T.inside) s0 = 0.f, s1 = 0.f, s2 = T.q.z, s3 = T.q.w, s4 = T.s;
    float r4 = (T.q.z + s2) * 0.5f;
    float r5 = (T.q.w + s3) * 0.5f;
    float r6 = (T.s + s4) * 0.25f;
    float r2 = (T.q.x - s0) * 0.5f;
    float r3 = (T.q.y - s1) * 0.5f;
    r2 += r4 * T.dy + r6 * T.dx;
    r3 += r6 * T.dy + r5 * T.dx;
    if (border && (gx < 5 || gx >= w - 5 || gy < 5 || gy >= h - 5)) {
        const float sc = (gx < 5 ? c_border[gx] : 1.f) * (gx >= w - 5 ? c_border[w - 1 - gx] : 1.f) *
                         (gy < 5 ? c_border[gy] : 1.f) * (gy >= h - 5 ? c_border[h - 1 - gy] : 1.f);
        r2 *= sc, r3 *= sc, r4 *= sc, r5 *= sc, r6 *= sc;
    }
    M[0] = r4 * r4 + r6 * r6;
    M[1] = (r4 + r5) * r6;
    M[2] = r5 * r5 + r6 * r6;
    M[3] = r4 * r2 + r6 * r3;
    M[4] = r6 * r2 + r5 * r3;
}

// L2 prefetch of everything the NEXT group's M phase will read (R0, R1 at zero displacement, flow):
// a group's columns are new to the whole GPU, so without this every gather of the M phase waits on
// DRAM; with it they wait on L2.  One 128-byte line per instruction, 12 lines per row; the row
// index runs fastest so a warp works on one kind of line (cp.async.bulk.prefetch was measured
// slower than no prefetch at all for these 128..512-byte runs).
template <typename T>
__device__ __forceinline__ void xm_prefetch(const float4* r0q, const float* r0s, const float4* r1q,
                                            const float* r1s, const float2* fb, int w, int h, int cx, int ry0) {
    if (cx >= w) return;
    for (int i = threadIdx.x; i < T::RH * 12; i += T::NT) {
        const int j = i / T::RH, row = i - j * T::RH;
        const int o = min(max(ry0 + row, 0), h - 1) * w + cx;
        const char* p;
        if (j < 4)
            p = reinterpret_cast<const char*>(r0q + o) + j * 128;
        else if (j < 8)
            p = reinterpret_cast<const char*>(r1q + o) + (j - 4) * 128;
        else if (j == 8)
            p = reinterpret_cast<const char*>(r0s + o);
        else if (j == 9)
            p = reinterpret_cast<const char*>(r1s + o);
        else
            p = reinterpret_cast<const char*>(fb + o) + (j - 10) * 128;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
}

// M for nc = 1 << ncl2 (32 or 8) columns starting at cx0, rows ry0 .. ry0 + RH - 1, into sM[c][row][col].
// One instance serves the lead-in, the groups and the tail (ptxas merges separate instances into
// one 160-register monster); coordinates are clamped unconditionally (replicate border).
template <typename T>
__device__ __forceinline__ void xm_m_phase(float* __restrict__ sM, const float4* r0q, const float* r0s,
                                           const float4* r1q, const float* r1s, const float2* fb, int w, int h,
                                           int cx0, int ry0, int ncl2) {
    constexpr int RH = T::RH, MS = T::MS, NW = T::NT / 32;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int col = lane & ((1 << ncl2) - 1), rsub = lane >> ncl2;
    const int rpw = 32 >> ncl2;  // rows one warp instruction covers
    const int gx = min(max(cx0 + col, 0), w - 1);
    const float xf = static_cast<float>(gx);
    const int step = NW * rpw;
    // the 5-pixel attenuation band can only be met by groups at the image edge
    const bool border = cx0 < 5 || cx0 + 32 > w - 5 || ry0 < 5 || ry0 + RH > h - 5;
    // PIX pixels (rows r, r + step, ..) per trip: all their gathers are issued before any is
    // consumed, and the next trip's flow vectors are already in flight.  Rows are clamped, so
    // every load is in bounds even past the last row; only the store is guarded.
    constexpr int PIX = T::PIX;
    const int r0 = wi * rpw + rsub;
    float* dst = sM + r0 * MS + col;
    // every trip's flow vectors are loaded up front; trips unrolled
    constexpr int TRIPS = (RH + PIX * NW - 1) / (PIX * NW);  // 32-column groups; the 8-column ones need one
    float2 f[TRIPS][PIX];
#pragma unroll
    for (int t = 0; t < TRIPS; ++t)
#pragma unroll
        for (int p = 0; p < PIX; ++p)
            f[t][p] = __ldg(at_index<8>(fb, min(max(ry0 + r0 + (t * PIX + p) * step, 0), h - 1) * w + gx));
#pragma unroll
    for (int t = 0; t < TRIPS; ++t) {
        const int r = r0 + t * PIX * step;
        if (r >= RH) break;
        XmTaps Ta[PIX];
        int gy[PIX];
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
            gy[p] = min(max(ry0 + r + p * step, 0), h - 1);
            xm_gather(r0q, r0s, r1q, r1s, w, h, xf, gx, gy[p], f[t][p], Ta[p]);
        }
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
            float Mv[5];
            xm_finish(Ta[p], border, w, h, gx, gy[p], Mv);
            if (p == 0 || r + p * step < RH) {
#pragma unroll
                for (int c = 0; c < 5; ++c) dst[c * RH * MS + (t * PIX + p) * step * MS] = Mv[c];
            }
        }
    }
}

// vertical sums of nc new columns -> window columns vpos0 .. vpos0 + nc - 1
template <typename T>
__device__ __forceinline__ void xm_v_phase(const float* sM, float* sV, int vpos0, int ncl2, int tid) {
    constexpr int PART = T::TY / T::VSPLIT;
    const int ntask = (5 * T::VSPLIT) << ncl2;
    for (int i = tid; i < ntask; i += T::NT) {
        const int col = i & ((1 << ncl2) - 1), t = i >> ncl2;
        const int part = t % T::VSPLIT, c = t / T::VSPLIT;
        window_sums<T::WIN, PART>(sM + (c * T::RH + part * PART) * T::MS + col, T::MS,
                                  sV + (c * T::TY + part * PART) * T::VS + vpos0 + col, T::VS);
    }
}

// horizontal sums over the 46-column window, in place: the 32 finished columns land on window
// columns 14..45 (where the next group's vertical sums will be written after the solve has read
// them) and the last 14 window columns move to the front
// fill: columns outside the image are replicas of the edge column (replicate border), so their vertical sums are
// copies of the edge column's — the thread that owns a row writes them before it sums the row, instead of a
// lead-in / tail group evaluating M for them.  fill_n values from window column fill_src to fill_dst.. (0: none)
template <typename T>
__device__ __forceinline__ void xm_h_phase(float* sV, int tid, int fill_src = 0, int fill_dst = 0, int fill_n = 0) {
    for (int i = tid; i < 5 * T::TY; i += T::NT) {   // i = c * TY + row
        float* v = sV + i * T::VS;
        if (fill_n) {
            const float e = v[fill_src];
            for (int j = 0; j < fill_n; ++j) v[fill_dst + j] = e;
        }
        window_sums_carry<T::WIN, 32>(v);
    }
}

// 2x2 solve on the raw window sums: norm^2 scales the determinant and both numerators alike
// (the regulariser 1e-3 belongs to the normalised determinant, as in solve_flow)
template <typename T>
__device__ __forceinline__ void xm_solve(const float* sV, float2* __restrict__ fo, int w, int h, int y0, int c0,
                                         int xlo, int xhi, float norm, int tid) {
    const int lane = tid & 31, wi = tid >> 5;
    const int gx = c0 - T::HM + lane;
    if (gx < xlo || gx >= xhi) return;
    const int rows = min(T::TY, h - y0);
    const float n2 = norm * norm;
    float2* out = fo + static_cast<size_t>(y0 + wi) * w + gx;
    const float* g = sV + wi * T::VS + (T::WIN - 1) + lane;
    for (int row = wi; row < rows; row += T::NT / 32, out += (T::NT / 32) * w, g += (T::NT / 32) * T::VS) {
        const float g0 = g[0], g1 = g[T::TY * T::VS], g2 = g[2 * T::TY * T::VS], g3 = g[3 * T::TY * T::VS],
                    g4 = g[4 * T::TY * T::VS];
        const float t = n2 * __fdividef(1.f, fmaf(g0 * g2 - g1 * g1, n2, 1e-3f));
        *out = make_float2((g0 * g4 - g1 * g3) * t, (g2 * g3 - g1 * g4) * t);
    }
}

template <typename T>
__global__ void __launch_bounds__(T::NT, T::MINB) k_flow_iter_xm(const float* __restrict__ R0,
                                                                 const float* __restrict__ R1,
                                                                 const float2* __restrict__ flow_in,
                                                                 float2* __restrict__ flow_out, int w, int h, int seg,
                                                                 float norm, int prefetch, int edge_fill) {
    extern __shared__ float sM[];     // [5][RH][MS]
    float* sV = sM + T::M_FLOATS;     // [5][TY][VS]
    const int x0 = blockIdx.x * seg, y0 = blockIdx.y * T::TY, b = blockIdx.z;
    const int xhi = min(x0 + seg, w);
    const int ngroups = (xhi - x0 + 31) >> 5;
    const int ry0 = y0 - T::HM;
    const size_t plane = static_cast<size_t>(w) * h;
    const RView R0b = r_view(R0, gridDim.z, b, plane), R1b = r_view(R1, gridDim.z, b, plane);
    const float4* r0q = R0b.q;
    const float* r0s = R0b.s;
    const float4* r1q = R1b.q;
    const float* r1s = R1b.s;
    const float2* fb = flow_in + static_cast<size_t>(b) * plane;
    float2* fo = flow_out + static_cast<size_t>(b) * plane;
    // k = -1: lead-in, the 8 columns left of the segment (only the last 7 are read) -> window columns
    // 6..13; k = ngroups: tail, the 8 columns right of it that finish its last 7 outputs.  At the image's
    // left / right edge those columns are replicas of column 0 / w - 1: no M, no vertical sums — the H phase
    // copies the edge column's sums (bit-identical: the replicas' M values are the edge column's).
    for (int k = -1; k <= ngroups; ++k) {
        const bool lead = k < 0, tail = k == ngroups;
        const int c0 = lead ? x0 - 8 : x0 + 32 * k;
        if (tail && c0 - T::HM >= xhi) break;
        if (lead && x0 == 0 && edge_fill) continue;
        const int ncl2 = (lead || tail) ? 3 : 5;
        // window columns 0..13 hold columns c0 - 14 .. c0 - 1; at the right edge column w - 1 is one of the last 7
        const bool rep_tail = tail && c0 >= w && edge_fill;
        if (!rep_tail) {
            // M(k) may start as soon as V(k-1) has left sM: the barrier in front of H(k-1) saw to that,
            // so the solve of group k-1 and this M phase share one barrier interval
            if (prefetch && k + 1 < ngroups) xm_prefetch<T>(r0q, r0s, r1q, r1s, fb, w, h, x0 + 32 * (k + 1), ry0);
            xm_m_phase<T>(sM, r0q, r0s, r1q, r1s, fb, w, h, c0, ry0, ncl2);
            __syncthreads();  // M complete; S(k-1) has read the window columns V(k) overwrites
            xm_v_phase<T>(sM, sV, lead ? 6 : 14, ncl2, threadIdx.x);
        }
        __syncthreads();      // (replica tail: S(k-1) has read the window columns the fill overwrites)
        if (lead) continue;
        int fill_src = 0, fill_dst = 0, fill_n = 0;
        if (rep_tail)
            fill_src = w + 13 - c0, fill_dst = 14, fill_n = 8;      // columns w .. = column w - 1
        else if (k == 0 && x0 == 0 && edge_fill)
            fill_src = 14, fill_dst = 7, fill_n = 7;                // columns -7 .. -1 = column 0
        xm_h_phase<T>(sV, threadIdx.x, fill_src, fill_dst, fill_n);
        __syncthreads();
        xm_solve<T>(sV, fo, w, h, y0, c0, x0, xhi, norm, threadIdx.x);
    }
}

size_t flow_iter_smem(int m) {
    int RW = FI_TX + 2 * m, RH = FI_TY + 2 * m, SW = RW | 1;
    return static_cast<size_t>(5) * (RH + FI_TY) * SW * sizeof(float);
}

// ------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------
// img1 != nullptr: a second array of B images handled by the same two launches (T and out then hold
// 2 B images: the caller checks that T is large enough)
int launch_pyr(datmo_ctx* h, const void* img, const void* img1, int dtype, int H, int W, int B, const FbLayer& L,
               const float* d_tab, float* T, float* out) {
    const double* hf = reinterpret_cast<const double*>(d_tab);
    const double* vf = hf + L.w;
    const int* hx = reinterpret_cast<const int*>(vf + L.h);
    const int* vy = hx + L.w;
    const float* gk = reinterpret_cast<const float*>(vy + L.h);
    const int pr = L.ksize >> 1;
    const size_t smem = static_cast<size_t>(PYR_ROWS) * (((pr + 3) & ~3) + ((W + pr + 1 + 3) & ~3)) * sizeof(float);
    DATMO_REQUIRE(h, smem <= 227 * 1024, "image too wide for the pyramid row staging");
    static SmemGrant grant_u8, grant_f32;
    const int nimg = img1 ? 2 * B : B;
    static const bool h_cols = getenv("DATMO_PYR_H_COLS") != nullptr;   // A/B runs against k_pyr_h
    const size_t smem_rows = static_cast<size_t>(PH2_ROWS) * 4 * (((((pr + 3) & ~3) + W + pr + 1 + 3) >> 2) | 1) +
                             static_cast<size_t>(PH2_ROWS) * (L.w | 1) * sizeof(float);
    PyrTaps taps;
    {
        const std::vector<float> g = gaussian_kernel(L.ksize, L.sigma);
        for (int i = 0; i < PH2_MAX_K; ++i) taps.k[i] = i < L.ksize ? g[i] : 0.f;
    }
    if (dtype == DATMO_U8 && !h_cols && L.ksize <= PH2_MAX_K && smem_rows <= 200 * 1024) {
        const int vec4 = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(img) & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(img1) & 3) == 0;
        dim3 g1(ceil_div(H, PH2_ROWS), nimg);
        static const int h_threads =
            getenv("DATMO_PYR_H_THREADS") ? std::min(PH2_THREADS, std::max(32, atoi(getenv("DATMO_PYR_H_THREADS")) & ~31))
                                          : PH2_THREADS;
        static SmemGrant grant7, grant25, grant0;
        auto go = [&](auto kern, SmemGrant& grant) -> int {
            DATMO_TRY(datmo_grant_smem(h, kern, smem_rows, grant));
            LaunchScope ls(h, DATMO_TAG_PYRAMID);
            kern<<<g1, h_threads, smem_rows, h->stream>>>(static_cast<const uint8_t*>(img),
                                                            static_cast<const uint8_t*>(img1), B, T, H, W, L.w, taps,
                                                            L.ksize, hx, hf, vec4);
            return DATMO_OK;
        };
        if (L.ksize == 7) DATMO_TRY(go(k_pyr_h_rows<7>, grant7));
        else if (L.ksize == 25) DATMO_TRY(go(k_pyr_h_rows<25>, grant25));
        else DATMO_TRY(go(k_pyr_h_rows<0>, grant0));
        DATMO_POST_LAUNCH(h);
    } else {
    dim3 g1(ceil_div(H, PYR_ROWS), nimg);
    const int vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(img1) & 15) == 0;
    {
        if (dtype == DATMO_U8) {
            DATMO_TRY(datmo_grant_smem(h, k_pyr_h<uint8_t>, smem, grant_u8));
            LaunchScope ls(h, DATMO_TAG_PYRAMID);
            k_pyr_h<uint8_t><<<g1, 256, smem, h->stream>>>(static_cast<const uint8_t*>(img),
                                                           static_cast<const uint8_t*>(img1), B, T, H, W, L.w, gk,
                                                           L.ksize, hx, hf, vec);
        } else {
            DATMO_TRY(datmo_grant_smem(h, k_pyr_h<float>, smem, grant_f32));
            LaunchScope ls(h, DATMO_TAG_PYRAMID);
            k_pyr_h<float><<<g1, 256, smem, h->stream>>>(static_cast<const float*>(img), static_cast<const float*>(img1),
                                                         B, T, H, W, L.w, gk, L.ksize, hx, hf, vec);
        }
    }
    DATMO_POST_LAUNCH(h);
    }
    dim3 g2(ceil_div(L.w, 128), ceil_div(L.h, PYR_VROWS), nimg);
    static const bool v_old = getenv("DATMO_PYR_V_OLD") != nullptr;   // A/B runs against k_pyr_v
    {
        LaunchScope ls(h, DATMO_TAG_PYRAMID);
        if (v_old || L.ksize > PH2_MAX_K)
            k_pyr_v<<<g2, 128, 0, h->stream>>>(T, out, H, L.w, L.h, gk, L.ksize, vy, vf);
        else if (L.ksize == 7)
            k_pyr_v_taps<7><<<g2, 128, 0, h->stream>>>(T, out, H, L.w, L.h, taps, L.ksize, vy, vf);
        else if (L.ksize == 25)
            k_pyr_v_taps<25><<<g2, 128, 0, h->stream>>>(T, out, H, L.w, L.h, taps, L.ksize, vy, vf);
        else
            k_pyr_v_taps<0><<<g2, 128, 0, h->stream>>>(T, out, H, L.w, L.h, taps, L.ksize, vy, vf);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

// coarse layers through the finest layer's kernel without its blur stage (packed arithmetic, bulk-copy row stores)
template <int N>
int launch_polyexp_t(datmo_ctx* h, const float* I, float* R, int w, int hh, int B, int imgs_per_array,
                     const PolyCoef& pc) {
    constexpr int MINB = N <= 5 ? 4 : 2;
    static SmemGrant grant;
    DATMO_TRY(datmo_grant_smem(h, k_pyr0_polyexp_t<float, N, MINB, false>, P0T<N>::SMEM, grant));
    dim3 g(ceil_div(w, P0_TX), ceil_div(hh, P0T<N>::TY), B);
    {
        LaunchScope ls(h, DATMO_TAG_POLYEXP);
        k_pyr0_polyexp_t<float, N, MINB, false><<<g, P0_THREADS, P0T<N>::SMEM, h->stream>>>(I, nullptr, B, R, w, hh,
                                                                                          imgs_per_array, pc);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

int launch_polyexp(datmo_ctx* h, const float* I, float* R, int w, int hh, int B, int imgs_per_array,
                   const PolyCoef& pc) {
    static const bool generic = getenv("DATMO_POLYEXP_GENERIC") != nullptr;   // A/B runs against k_polyexp
    if (!generic && (reinterpret_cast<uintptr_t>(R) & 15) == 0) {
        if (pc.n == 5) return launch_polyexp_t<5>(h, I, R, w, hh, B, imgs_per_array, pc);
        if (pc.n == 7) return launch_polyexp_t<7>(h, I, R, w, hh, B, imgs_per_array, pc);
    }
    int RW = PE_TX + 2 * pc.n, RH = PE_TY + 2 * pc.n;
    size_t smem = static_cast<size_t>(RH * RW + 3 * PE_TY * RW) * sizeof(float);
    static SmemGrant grant;
    DATMO_TRY(datmo_grant_smem(h, k_polyexp, smem, grant));
    dim3 g(ceil_div(w, PE_TX), ceil_div(hh, PE_TY), B);
    {
        LaunchScope ls(h, DATMO_TAG_POLYEXP);
        k_polyexp<<<g, PE_THREADS, smem, h->stream>>>(I, R, w, hh, imgs_per_array, pc);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

int flow_tile_choice() {
    const char* e = getenv("DATMO_FI_TILE");
    return e ? atoi(e) : DATMO_FI_TILE_DEFAULT;
}

// finest layer: fused blur + polyexp when the tap count has a compile-time instance
bool pyr0_polyexp_supported(int poly_n) { return (poly_n == 5 || poly_n == 7) && !getenv("DATMO_NO_PYR0_FUSION"); }

// prev and next frames of the batch in one launch, output through the bulk-copy engine; needs W % 4 == 0
// and 16-byte aligned frames.  R: the two consecutive R arrays (R0 | R1) of B images each.
bool pyr0_bulk_supported(const void* prev, const void* next, int dtype, int W, int poly_n) {
    if (getenv("DATMO_PYR0_LSU")) return false;   // A/B runs against k_pyr0_polyexp
    const size_t align = dtype == DATMO_U8 ? 4 : 16;
    return (poly_n == 5 || poly_n == 7) && (W & 3) == 0 && reinterpret_cast<uintptr_t>(prev) % align == 0 &&
           reinterpret_cast<uintptr_t>(next) % align == 0;
}

template <typename SrcT, int N>
int launch_pyr0_bulk_t(datmo_ctx* h, const void* prev, const void* next, int H, int W, int B, float* R,
                       const PolyCoef& pc) {
    // CTAs per SM the register budget is set for: 4 (64 registers, a few spilled words) against 3 (78)
    static const int minb = getenv("DATMO_PYR0_MINB") ? atoi(getenv("DATMO_PYR0_MINB")) : (N <= 5 ? 4 : 2);
    static SmemGrant grant4, grant3;
    dim3 g(ceil_div(W, P0_TX), ceil_div(H, P0T<N>::TY), 2 * B);
    if (minb >= 4 && N <= 5) {
        DATMO_TRY(datmo_grant_smem(h, k_pyr0_polyexp_t<SrcT, N, 4>, P0T<N>::SMEM, grant4));
        LaunchScope ls(h, DATMO_TAG_POLYEXP);
        k_pyr0_polyexp_t<SrcT, N, 4><<<g, P0_THREADS, P0T<N>::SMEM, h->stream>>>(
            static_cast<const SrcT*>(prev), static_cast<const SrcT*>(next), B, R, W, H, B, pc);
    } else {
        DATMO_TRY(datmo_grant_smem(h, k_pyr0_polyexp_t<SrcT, N, 2>, P0T<N>::SMEM, grant3));
        LaunchScope ls(h, DATMO_TAG_POLYEXP);
        k_pyr0_polyexp_t<SrcT, N, 2><<<g, P0_THREADS, P0T<N>::SMEM, h->stream>>>(
            static_cast<const SrcT*>(prev), static_cast<const SrcT*>(next), B, R, W, H, B, pc);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

int launch_pyr0_bulk(datmo_ctx* h, const void* prev, const void* next, int dtype, int H, int W, int B, float* R,
                     const PolyCoef& pc) {
    if (dtype == DATMO_U8)
        return pc.n == 5 ? launch_pyr0_bulk_t<uint8_t, 5>(h, prev, next, H, W, B, R, pc)
                         : launch_pyr0_bulk_t<uint8_t, 7>(h, prev, next, H, W, B, R, pc);
    return pc.n == 5 ? launch_pyr0_bulk_t<float, 5>(h, prev, next, H, W, B, R, pc)
                     : launch_pyr0_bulk_t<float, 7>(h, prev, next, H, W, B, R, pc);
}

int launch_pyr0_polyexp(datmo_ctx* h, const void* img, int dtype, int H, int W, int B, float* R, const PolyCoef& pc) {
    dim3 g(ceil_div(W, P0_TX), ceil_div(H, pc.n <= 5 ? P0Tile<5>::TY : P0Tile<7>::TY), B);
    {
        LaunchScope ls(h, DATMO_TAG_POLYEXP);
        if (dtype == DATMO_U8) {
            if (pc.n == 5)
                k_pyr0_polyexp<uint8_t, 5><<<g, P0_THREADS, 0, h->stream>>>(static_cast<const uint8_t*>(img), R, W, H, B, pc);
            else
                k_pyr0_polyexp<uint8_t, 7><<<g, P0_THREADS, 0, h->stream>>>(static_cast<const uint8_t*>(img), R, W, H, B, pc);
        } else {
            if (pc.n == 5)
                k_pyr0_polyexp<float, 5><<<g, P0_THREADS, 0, h->stream>>>(static_cast<const float*>(img), R, W, H, B, pc);
            else
                k_pyr0_polyexp<float, 7><<<g, P0_THREADS, 0, h->stream>>>(static_cast<const float*>(img), R, W, H, B, pc);
        }
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

template <int TX, int TY, int NT, int MINB, bool FUSED>
int launch_flow_iter_w(datmo_ctx* h, const float* R0, const float* R1, const float* flow_in, const float* Min,
                       float* flow_out, int w, int hh, int B, float norm) {
    using T = FlowTile<TX, TY, 15, NT>;
    static SmemGrant grant;
    DATMO_TRY(datmo_grant_smem(h, k_flow_iter_w<TX, TY, 15, NT, MINB, FUSED>, T::SMEM, grant));
    dim3 g(ceil_div(w, TX), ceil_div(hh, TY), B);
    {
        LaunchScope ls(h, DATMO_TAG_FLOW_ITER);
        k_flow_iter_w<TX, TY, 15, NT, MINB, FUSED><<<g, T::THREADS, T::SMEM, h->stream>>>(
            R0, R1, reinterpret_cast<const float2*>(flow_in), Min, reinterpret_cast<float2*>(flow_out), w, hh, norm);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

// Band height and segment length for the x-marching kernel: the pair that minimises
// (CTA waves, rounded up) x (columns evaluated per segment, lead-in and tail included) x (M rows per band).
// The finest layer of a large batch runs many waves and wants the tall band (least halo); a coarse layer
// that fits one wave wants the band height that just fills the CTA slots.
// The band height decides where the vertical window sums split their 15-row blocks, i.e. the rounding of the
// sums, so it must depend on the layer's GEOMETRY only — a pair's flow may not change with the batch it is
// computed in (tests: batched == single, shard equality).  It is therefore planned for a nominal batch of 32
// on a 148-SM part; the segment length (no effect on the arithmetic: groups stay aligned to multiples of 32
// columns) is planned for the actual batch and device.
struct XmPlan {
    int ty, seg;
};
constexpr int XM_TY_TALL = 46, XM_TY_SHORT = 36;

XmPlan xm_plan_for(int w, int hh, int B, int slots, int only_ty) {
    static const int seg_env = getenv("DATMO_XM_SEG") ? atoi(getenv("DATMO_XM_SEG")) : 0;
    XmPlan best{XM_TY_TALL, 32};
    double best_cost = 1e300;
    for (int ty : {XM_TY_TALL, XM_TY_SHORT}) {
        if (only_ty && ty != only_ty) continue;
        const int bands = ceil_div(hh, ty);
        for (int seg = 32; seg <= ((w + 31) & ~31); seg += 32) {
            if (seg_env > 0 && seg != ((seg_env + 31) & ~31)) continue;
            const int nseg = ceil_div(w, seg);
            const double ctas = static_cast<double>(nseg) * bands * B;
            const double waves = ceil(ctas / slots);
            // + ~24 columns' worth of per-CTA fixed cost
            const double cost = waves * (seg + 16 + 24) * (ty + 14);
            if (cost < best_cost - 1e-9) best_cost = cost, best = XmPlan{ty, seg};
        }
    }
    return best;
}

XmPlan xm_plan(int w, int hh, int B, int slots) {
    static const int ty_env = getenv("DATMO_XM_TY") ? atoi(getenv("DATMO_XM_TY")) : 0;
    const int ty = ty_env ? ty_env : xm_plan_for(w, hh, 32, 2 * 148, 0).ty;
    return xm_plan_for(w, hh, B, slots, ty);
}

template <typename T>
int launch_flow_iter_xm(datmo_ctx* h, const float* R0, const float* R1, const float* flow_in, float* flow_out, int w,
                        int hh, int B, float norm, int seg) {
    static SmemGrant grant;
    DATMO_TRY(datmo_grant_smem(h, k_flow_iter_xm<T>, T::SMEM, grant));
    static const int prefetch = getenv("DATMO_XM_PREFETCH") ? atoi(getenv("DATMO_XM_PREFETCH")) : 1;
    static const int edge_fill = getenv("DATMO_XM_EDGE_FILL") ? atoi(getenv("DATMO_XM_EDGE_FILL")) : 1;   // A/B switch
    dim3 g(ceil_div(w, seg), ceil_div(hh, T::TY), B);
    {
        LaunchScope ls(h, DATMO_TAG_FLOW_ITER);
        k_flow_iter_xm<T><<<g, T::NT, T::SMEM, h->stream>>>(R0, R1, reinterpret_cast<const float2*>(flow_in),
                                                           reinterpret_cast<float2*>(flow_out), w, hh, seg, norm,
                                                           prefetch, edge_fill);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

template <bool FUSED>
int launch_flow_iter(datmo_ctx* h, const float* R0, const float* R1, const float* flow_in, const float* Min,
                     float* flow_out, int w, int hh, int B, int winsize) {
    int m = winsize / 2;
    float norm = static_cast<float>(1.0 / (static_cast<double>(winsize) * winsize));
    if (m == 7 && !getenv("DATMO_GENERIC_FLOW_ITER")) {
        // the reference's winsize 15 (and 14): compile-time window.  The fused iteration runs the
        // x-marching kernel; DATMO_FI_TILE=0 selects the 64x32 tile kernel it replaced (A/B runs,
        // DESIGN.md §5), which also serves the blur-and-solve-only entry point.
        static const int tile = flow_tile_choice();
        if (FUSED && tile != 0) {
            const XmPlan plan = xm_plan(w, hh, B, h->sm_count * 2);
            if (plan.ty == XM_TY_SHORT)
                return launch_flow_iter_xm<XmTile<XM_TY_SHORT, 320, 2, 2, 2>>(h, R0, R1, flow_in, flow_out, w, hh, B,
                                                                              norm, plan.seg);
            return launch_flow_iter_xm<XmTile<XM_TY_TALL, 320, 2, 2, 2>>(h, R0, R1, flow_in, flow_out, w, hh, B, norm,
                                                                         plan.seg);
        }
        return launch_flow_iter_w<64, 32, 256, 2, FUSED>(h, R0, R1, flow_in, Min, flow_out, w, hh, B, norm);
    }
    size_t smem = flow_iter_smem(m);
    DATMO_REQUIRE(h, smem <= 227 * 1024, "winsize too large for the flow-iteration tile");
    static SmemGrant grant;
    DATMO_TRY(datmo_grant_smem(h, k_flow_iter<FUSED>, smem, grant));
    dim3 g(ceil_div(w, FI_TX), ceil_div(hh, FI_TY), B);
    {
        LaunchScope ls(h, DATMO_TAG_FLOW_ITER);
        k_flow_iter<FUSED><<<g, FI_THREADS, smem, h->stream>>>(R0, R1, reinterpret_cast<const float2*>(flow_in), Min,
                                                               reinterpret_cast<float2*>(flow_out), w, hh, m, norm);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

int launch_update_matrices(datmo_ctx* h, const float* R0, const float* R1, const float* flow, float* M, int w,
                           int hh, int B) {
    dim3 g(ceil_div(w, 256), hh, B);
    {
        LaunchScope ls(h, DATMO_TAG_FLOW_ITER);
        k_update_matrices<<<g, 256, 0, h->stream>>>(R0, R1, reinterpret_cast<const float2*>(flow), M, w, hh);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

int launch_upsample(datmo_ctx* h, const float* fin, int hi, int wi, float* fout, int ho, int wo, int B, double mul,
                    const float* d_tab) {
    const int* sx = reinterpret_cast<const int*>(d_tab);
    const float* fx = d_tab + wo;
    const int* sy = reinterpret_cast<const int*>(fx + wo);
    const float* fy = fx + wo + ho;
    dim3 g(ceil_div(wo, 512), ceil_div(ho, UP_ROWS), B);
    {
        LaunchScope ls(h, DATMO_TAG_FLOW_INIT);
        k_upsample_flow<<<g, 256, 0, h->stream>>>(reinterpret_cast<const float2*>(fin), reinterpret_cast<float2*>(fout),
                                                  hi, wi, ho, wo, sx, fx, sy, fy, static_cast<float>(mul));
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

// planar [B][5][n] <-> the library's R layout ([B][n] float4 | [B][n] float); only the
// stage-level entry points the parity tests use go through these
__global__ void __launch_bounds__(256) k_planar_to_r(const float* __restrict__ P, float* __restrict__ R, int B,
                                                     size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= n) return;
    const float* p = P + static_cast<size_t>(b) * 5 * n + i;
    float4* q;
    float* sdst;
    r_out(R, B, b, n, q, sdst);
    q[i] = make_float4(p[0], p[n], p[2 * n], p[3 * n]);
    sdst[i] = p[4 * n];
}

__global__ void __launch_bounds__(256) k_r_to_planar(const float* __restrict__ R, float* __restrict__ P, int B,
                                                     size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= n) return;
    const RView v = r_view(R, B, b, n);
    const float4 q = v.q[i];
    float* p = P + static_cast<size_t>(b) * 5 * n + i;
    p[0] = q.x, p[n] = q.y, p[2 * n] = q.z, p[3 * n] = q.w, p[4 * n] = v.s[i];
}

int launch_planar_to_r(datmo_ctx* h, const float* P, float* R, int B, size_t n) {
    dim3 g(static_cast<unsigned>((n + 255) / 256), B);
    {
        LaunchScope ls(h, DATMO_TAG_POLYEXP);
        k_planar_to_r<<<g, 256, 0, h->stream>>>(P, R, B, n);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

int launch_r_to_planar(datmo_ctx* h, const float* R, float* P, int B, size_t n) {
    dim3 g(static_cast<unsigned>((n + 255) / 256), B);
    {
        LaunchScope ls(h, DATMO_TAG_POLYEXP);
        k_r_to_planar<<<g, 256, 0, h->stream>>>(R, P, B, n);
    }
    DATMO_POST_LAUNCH(h);
    return DATMO_OK;
}

int check_params(datmo_ctx* h, int H, int W, int B, const datmo_farneback_params* p) {
    DATMO_REQUIRE(h, p != nullptr, "params is null");
    DATMO_REQUIRE(h, H >= 2 && W >= 2 && B >= 1, "need H, W >= 2 and batch >= 1");
    DATMO_REQUIRE(h, H <= 65535 && B <= 65535, "H and batch must fit a CUDA grid dimension");
    DATMO_REQUIRE(h, p->flags == 0, "only flags = 0 is on the reference path (main.py:139)");
    DATMO_REQUIRE(h, p->pyr_scale > 0 && p->pyr_scale < 1, "pyr_scale must be in (0, 1)");
    DATMO_REQUIRE(h, p->levels >= 1 && p->winsize >= 1 && p->iterations >= 1, "levels, winsize, iterations >= 1");
    DATMO_REQUIRE(h, p->poly_n >= 1 && p->poly_n <= POLY_MAX_N, "poly_n out of range");
    DATMO_REQUIRE(h, flow_iter_smem(p->winsize / 2) <= 227 * 1024, "winsize too large");
    return DATMO_OK;
}

struct FbWorkspace {
    float* kern;   // gaussian taps of every layer, concatenated
    float* T;      // [B][H][wmax], reused for prev then next
    float* I;      // [2][B][h][w]
    float* R;      // [2][B][5][h][w]
    float* flowA;  // [B][h][w][2]
    float* flowB;
    float* M;      // [B][5][h][w] (unfused variant only)
};

size_t fb_carve(Bump& bump, FbWorkspace& ws, int H, int W, int B, int n_kern, bool need_M) {
    size_t N0 = static_cast<size_t>(H) * W;
    (void)n_kern;  // the tables live outside the arena (datmo_ctx::fb_tab)
    ws.kern = nullptr;
    ws.T = bump.take<float>(B * N0);
    ws.I = bump.take<float>(2 * B * N0);
    ws.R = bump.take<float>(2 * r_array_stride(B, N0));
    ws.flowA = bump.take<float>(2 * B * N0);
    ws.flowB = bump.take<float>(2 * B * N0);
    ws.M = need_M ? bump.take<float>(5 * B * N0) : nullptr;
    return bump.off;
}

int fb_run_chunk(datmo_ctx* h, const void* prev, const void* next, int dtype, int H, int W, int B,
                 const datmo_farneback_params& p, const std::vector<FbLayer>& layers, const PolyCoef& pc,
                 const FbWorkspace& ws, const std::vector<int>& kern_off, const std::vector<int>& up_off,
                 float* flow_out) {
    float* cur = nullptr;  // flow of the layer just finished
    int cur_w = 0, cur_h = 0;
    const bool fused = p.variant != 1;
    for (size_t li = 0; li < layers.size(); ++li) {
        const FbLayer& L = layers[li];
        const size_t n = static_cast<size_t>(L.w) * L.h;
        const bool last_layer = li + 1 == layers.size();
        float* I0 = ws.I;
        float* I1 = ws.I + B * n;
        float* R0 = ws.R;
        float* R1 = ws.R + r_array_stride(B, n);
        if (L.k == 0 && L.w == W && L.h == H && pyr0_polyexp_supported(pc.n)) {
            if (pyr0_bulk_supported(prev, next, dtype, W, pc.n)) {
                DATMO_TRY(launch_pyr0_bulk(h, prev, next, dtype, H, W, B, R0, pc));   // R1 follows R0 at r_array_stride
            } else {
                DATMO_TRY(launch_pyr0_polyexp(h, prev, dtype, H, W, B, R0, pc));
                DATMO_TRY(launch_pyr0_polyexp(h, next, dtype, H, W, B, R1, pc));
            }
        } else {
            if (2 * L.w <= W) {
                // both frames in one pair of launches: T ([B][H][W] floats) holds 2 B images of width w <= W / 2,
                // and I1 follows I0 in memory
                DATMO_TRY(launch_pyr(h, prev, next, dtype, H, W, B, L, ws.kern + kern_off[li], ws.T, I0));
            } else {
                DATMO_TRY(launch_pyr(h, prev, nullptr, dtype, H, W, B, L, ws.kern + kern_off[li], ws.T, I0));
                DATMO_TRY(launch_pyr(h, next, nullptr, dtype, H, W, B, L, ws.kern + kern_off[li], ws.T, I1));
            }
            DATMO_TRY(launch_polyexp(h, ws.I, ws.R, L.w, L.h, 2 * B, B, pc));
        }
        float* fin;
        float* fout;
        if (cur == nullptr) {
            fin = ws.flowA;
            fout = ws.flowB;
            DATMO_CHECK_CUDA(h, cudaMemsetAsync(fin, 0, B * n * 2 * sizeof(float), h->stream));
        } else {
            fin = cur == ws.flowA ? ws.flowB : ws.flowA;
            fout = cur;
            DATMO_TRY(launch_upsample(h, cur, cur_h, cur_w, fin, L.h, L.w, B, 1.0 / p.pyr_scale, ws.kern + up_off[li]));
        }
        for (int it = 0; it < p.iterations; ++it) {
            float* dst = (last_layer && it == p.iterations - 1) ? flow_out : fout;
            if (fused) {
                DATMO_TRY(launch_flow_iter<true>(h, R0, R1, fin, nullptr, dst, L.w, L.h, B, p.winsize));
            } else {
                DATMO_TRY(launch_update_matrices(h, R0, R1, fin, ws.M, L.w, L.h, B));
                DATMO_TRY(launch_flow_iter<false>(h, R0, R1, fin, ws.M, dst, L.w, L.h, B, p.winsize));
            }
            std::swap(fin, fout);
        }
        cur = fin;  // after the swap, fin is what was just written
        cur_w = L.w;
        cur_h = L.h;
    }
    return DATMO_OK;
}

size_t ws_budget_bytes() {
    const char* e = getenv("DATMO_WS_BUDGET_MB");
    size_t mb = e ? strtoull(e, nullptr, 10) : 8192;
    if (mb < 64) mb = 64;
    return mb << 20;
}

int fb_run(datmo_ctx* h, const void* prev, const void* next, int dtype, int H, int W, int batch,
           const datmo_farneback_params* p, float* flow) {
    DATMO_TRY(check_params(h, H, W, batch, p));
    DATMO_REQUIRE(h, prev && next && flow, "null image / flow pointer");
    DATMO_REQUIRE(h, dtype == DATMO_U8 || dtype == DATMO_F32, "dtype must be DATMO_U8 or DATMO_F32");
    std::vector<FbLayer> layers = fb_plan(H, W, p->pyr_scale, p->levels);
    PolyCoef pc;
    DATMO_REQUIRE(h, poly_setup(p->poly_n, p->poly_sigma, pc), "polynomial expansion setup failed");
    std::vector<float> kern_all;  // per-layer pyramid tables (pyr_tables), concatenated
    std::vector<int> kern_off;
    for (auto& L : layers) {
        DATMO_REQUIRE(h, L.ksize / 2 < std::min(H, W), "image too small for the pyramid smoothing kernel");
        kern_off.push_back(static_cast<int>(kern_all.size()));
        pyr_tables(H, W, L, kern_all);
    }
    std::vector<int> up_off(layers.size(), 0);  // flow upsampling tables of the step INTO layer li
    for (size_t li = 1; li < layers.size(); ++li) {
        up_off[li] = static_cast<int>(kern_all.size());
        up_tables(layers[li - 1].h, layers[li - 1].w, layers[li].h, layers[li].w, kern_all);
        if (kern_all.size() & 1) kern_all.push_back(0.f);  // pyramid tables must start 8-byte aligned
    }
    const bool need_M = p->variant == 1;
    // chunk the batch so the workspace stays inside the budget
    FbWorkspace ws;
    size_t per_pair;
    {
        Bump dry(nullptr);
        per_pair = fb_carve(dry, ws, H, W, 1, static_cast<int>(kern_all.size()), need_M);
    }
    int chunk = static_cast<int>(std::max<size_t>(1, std::min<size_t>(batch, ws_budget_bytes() / per_pair)));
    chunk = std::min(chunk, 32767);  // polyexp runs 2 * chunk images in grid.z
    size_t total;
    {
        Bump dry(nullptr);
        total = fb_carve(dry, ws, H, W, chunk, static_cast<int>(kern_all.size()), need_M);
    }
    DATMO_TRY(datmo_ws_reserve(h, total));
    Bump bump(h->ws);
    fb_carve(bump, ws, H, W, chunk, static_cast<int>(kern_all.size()), need_M);
    // the tables live in their own device buffer and are uploaded only when the geometry changes
    if (h->fb_tab_host != kern_all) {
        DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));  // an earlier call may still read the old tables
        if (kern_all.size() > h->fb_tab_cap) {
            if (h->fb_tab) DATMO_CHECK_CUDA(h, cudaFree(h->fb_tab));
            h->fb_tab = nullptr;
            h->fb_tab_cap = 0;
            h->fb_tab_host.clear();
            DATMO_CHECK_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->fb_tab), kern_all.size() * sizeof(float)));
            h->fb_tab_cap = kern_all.size();
        }
        DATMO_CHECK_CUDA(h, cudaMemcpyAsync(h->fb_tab, kern_all.data(), kern_all.size() * sizeof(float),
                                            cudaMemcpyHostToDevice, h->stream));
        DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));  // kern_all is a stack-lifetime vector
        h->fb_tab_host = kern_all;
    }
    ws.kern = h->fb_tab;
    const size_t N0 = static_cast<size_t>(H) * W;
    const size_t esz = dtype == DATMO_U8 ? 1 : 4;
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        int B = std::min(chunk, batch - b0);
        DATMO_TRY(fb_run_chunk(h, static_cast<const char*>(prev) + b0 * N0 * esz,
                               static_cast<const char*>(next) + b0 * N0 * esz, dtype, H, W, B, *p, layers, pc, ws,
                               kern_off, up_off, flow + b0 * N0 * 2));
    }
    return DATMO_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
extern "C" {

void datmo_farneback_default_params(datmo_farneback_params* p) {
    if (!p) return;
    // Optical_flow/main.py:132-140
    p->pyr_scale = 0.3;
    p->levels = 5;
    p->winsize = 15;
    p->iterations = 5;
    p->poly_n = 5;
    p->poly_sigma = 5.0;
    p->flags = 0;
    p->variant = 0;
}

int datmo_farneback_layers(int H, int W, const datmo_farneback_params* p, int max_layers, int* w, int* hh) {
    if (!p || H < 1 || W < 1) return DATMO_E_INVALID;
    auto layers = fb_plan(H, W, p->pyr_scale, p->levels);
    for (size_t i = 0; i < layers.size() && static_cast<int>(i) < max_layers; ++i) {
        if (w) w[i] = layers[i].w;
        if (hh) hh[i] = layers[i].h;
    }
    return static_cast<int>(layers.size());
}

int datmo_farneback_dev(datmo_handle_t h, const void* prev, const void* next, int dtype, int H, int W, int batch,
                        const datmo_farneback_params* p, float* flow) {
    DATMO_ENTER(h);
    return fb_run(h, prev, next, dtype, H, W, batch, p, flow);
}

int datmo_farneback_host(datmo_handle_t h, const void* prev, const void* next, int dtype, int H, int W, int batch,
                         const datmo_farneback_params* p, float* flow) {
    DATMO_ENTER(h);
    DATMO_TRY(check_params(h, H, W, batch, p));
    DATMO_REQUIRE(h, prev && next && flow, "null image / flow pointer");
    DATMO_REQUIRE(h, dtype == DATMO_U8 || dtype == DATMO_F32, "dtype must be DATMO_U8 or DATMO_F32");
    const size_t N0 = static_cast<size_t>(H) * W;
    const size_t esz = dtype == DATMO_U8 ? 1 : 4;
    const size_t in_bytes = batch * N0 * esz, out_bytes = batch * N0 * 2 * sizeof(float);
    DATMO_TRY(datmo_io_reserve(h, 2 * in_bytes + out_bytes + 1024));   // grow-only: a steady stream of calls allocates once
    char* d_io = h->io;
    char* d_prev = d_io;
    char* d_next = d_io + ((in_bytes + 255) & ~size_t(255));
    float* d_flow = reinterpret_cast<float*>(d_next + ((in_bytes + 255) & ~size_t(255)));
    int st = DATMO_OK;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d_prev, prev, in_bytes, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_next, next, in_bytes, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        st = DATMO_E_CUDA;
    }
    if (st == DATMO_OK) st = fb_run(h, d_prev, d_next, dtype, H, W, batch, p, d_flow);
    if (st == DATMO_OK &&
        (e = cudaMemcpyAsync(flow, d_flow, out_bytes, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) {
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

int datmo_fb_pyramid_image_dev(datmo_handle_t h, const void* img, int dtype, int H, int W, int batch, int ksize,
                               double sigma, int h_out, int w_out, float* out) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, img && out && H >= 1 && W >= 1 && batch >= 1 && h_out >= 1 && w_out >= 1, "bad arguments");
    DATMO_REQUIRE(h, ksize >= 1 && (ksize & 1) && ksize / 2 < std::min(H, W), "bad ksize");
    FbLayer L{0, 1.0, sigma, ksize, w_out, h_out};
    std::vector<float> tab;
    pyr_tables(H, W, L, tab);
    Bump dry(nullptr);
    dry.take<float>(tab.size());
    dry.take<float>(static_cast<size_t>(batch) * H * w_out);
    DATMO_TRY(datmo_ws_reserve(h, dry.off));
    Bump bump(h->ws);
    float* d_k = bump.take<float>(tab.size());
    float* T = bump.take<float>(static_cast<size_t>(batch) * H * w_out);
    DATMO_CHECK_CUDA(h, cudaMemcpyAsync(d_k, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));  // tab is a stack-lifetime vector
    return launch_pyr(h, img, nullptr, dtype, H, W, batch, L, d_k, T, out);
}

int datmo_fb_polyexp_dev(datmo_handle_t h, const float* img, int hh, int ww, int batch, int poly_n, double poly_sigma,
                         float* R) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, img && R && hh >= 1 && ww >= 1 && batch >= 1, "bad arguments");
    PolyCoef pc;
    DATMO_REQUIRE(h, poly_setup(poly_n, poly_sigma, pc), "polynomial expansion setup failed");
    const size_t n = static_cast<size_t>(hh) * ww;
    DATMO_TRY(datmo_ws_reserve(h, 5 * batch * n * sizeof(float) + 256));
    float* tmp = reinterpret_cast<float*>(h->ws);
    DATMO_TRY(launch_polyexp(h, img, tmp, ww, hh, batch, batch, pc));
    return launch_r_to_planar(h, tmp, R, batch, n);   // the ABI hands out planar [batch][5][h][w]
}

int datmo_fb_update_matrices_dev(datmo_handle_t h, const float* R0, const float* R1, const float* flow, int hh, int ww,
                                 int batch, float* M) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, R0 && R1 && flow && M && hh >= 1 && ww >= 1 && batch >= 1, "bad arguments");
    const size_t n = static_cast<size_t>(hh) * ww;
    const size_t arr = (5 * batch * n * sizeof(float) + 255) & ~size_t(255);
    DATMO_TRY(datmo_ws_reserve(h, 2 * arr));
    float* r0 = reinterpret_cast<float*>(h->ws);
    float* r1 = reinterpret_cast<float*>(h->ws + arr);
    DATMO_TRY(launch_planar_to_r(h, R0, r0, batch, n));
    DATMO_TRY(launch_planar_to_r(h, R1, r1, batch, n));
    return launch_update_matrices(h, r0, r1, flow, M, ww, hh, batch);
}

int datmo_fb_blur_solve_dev(datmo_handle_t h, const float* M, int hh, int ww, int batch, int winsize, float* flow) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, M && flow && hh >= 1 && ww >= 1 && batch >= 1 && winsize >= 1, "bad arguments");
    return launch_flow_iter<false>(h, M, M, flow, M, flow, ww, hh, batch, winsize);
}

int datmo_fb_flow_iter_dev(datmo_handle_t h, const float* R0, const float* R1, const float* flow_in, int hh, int ww,
                           int batch, int winsize, float* flow_out) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, R0 && R1 && flow_in && flow_out && flow_in != flow_out, "bad arguments");
    DATMO_REQUIRE(h, hh >= 1 && ww >= 1 && batch >= 1 && winsize >= 1, "bad arguments");
    const size_t n = static_cast<size_t>(hh) * ww;
    const size_t arr = (5 * batch * n * sizeof(float) + 255) & ~size_t(255);
    DATMO_TRY(datmo_ws_reserve(h, 2 * arr));
    float* r0 = reinterpret_cast<float*>(h->ws);
    float* r1 = reinterpret_cast<float*>(h->ws + arr);
    DATMO_TRY(launch_planar_to_r(h, R0, r0, batch, n));
    DATMO_TRY(launch_planar_to_r(h, R1, r1, batch, n));
    return launch_flow_iter<true>(h, r0, r1, flow_in, nullptr, flow_out, ww, hh, batch, winsize);
}

int datmo_fb_upsample_flow_dev(datmo_handle_t h, const float* flow_in, int h_in, int w_in, int batch, int h_out,
                               int w_out, double mul, float* flow_out) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, flow_in && flow_out && h_in >= 1 && w_in >= 1 && h_out >= 1 && w_out >= 1 && batch >= 1,
                  "bad arguments");
    std::vector<float> tab;
    up_tables(h_in, w_in, h_out, w_out, tab);
    DATMO_TRY(datmo_ws_reserve(h, tab.size() * sizeof(float) + 256));
    float* d_tab = reinterpret_cast<float*>(h->ws);
    DATMO_CHECK_CUDA(h, cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));  // tab is a stack-lifetime vector
    return launch_upsample(h, flow_in, h_in, w_in, flow_out, h_out, w_out, batch, mul, d_tab);
}

}  // extern "C"
