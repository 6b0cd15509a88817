"""numpy restatement of ``cv2.calcOpticalFlowFarneback`` (flags=0) — TEST ORACLE.

The reference calls it at /root/reference/Optical_flow/main.py:142 with the
hard-coded parameters of main.py:132-140 (pyr_scale 0.3, levels 5, winsize 15,
iterations 5, poly_n 5, poly_sigma 5, flags 0).  The arithmetic lives in
opencv-python (unpinned by the reference; 4.13.0.92 in this image), file
``modules/video/src/optflowgf.cpp`` — not on disk, so this restates the
published algorithm step by step and is pinned
against the live wheel by ``tests/test_oracle_farneback.py`` and against
golden flows produced through the reference's own ``compute_velocity_vectors``.

Every step is exposed on its own so the CUDA kernels can be diffed stage by
stage.  Flow channel 0 = dx (columns), channel 1 = dy (rows).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64
FLT_EPSILON = 1.1920928955078125e-07
MIN_SIZE = 32
BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], dtype=F32)


def cv_round(x: float) -> int:
    """cvRound: round half to even (lrint)."""
    return int(np.rint(x))


# ----------------------------------------------------------------------------
# F0 / F1: which pyramid layers exist and their geometry
# ----------------------------------------------------------------------------
def level_plan(H: int, W: int, pyr_scale: float, levels: int):
    """Layers from coarsest to finest; each a dict(k, scale, sigma, ksize, w, h)."""
    k = 0
    scale = 1.0
    while k < levels:
        scale *= pyr_scale
        if W * scale < MIN_SIZE or H * scale < MIN_SIZE:
            break
        k += 1
    plan = []
    for kk in range(k, -1, -1):
        scale = 1.0
        for _ in range(kk):
            scale *= pyr_scale
        sigma = (1.0 / scale - 1.0) * 0.5
        ksize = max(cv_round(sigma * 5) | 1, 3)
        plan.append(dict(k=kk, scale=scale, sigma=sigma, ksize=ksize,
                         w=cv_round(W * scale), h=cv_round(H * scale)))
    return plan


# ----------------------------------------------------------------------------
# F2: pyramid image = GaussianBlur (full res, REFLECT_101) then bilinear resize
# ----------------------------------------------------------------------------
def gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    """cv::getGaussianKernel(ksize, sigma, CV_32F)."""
    if sigma <= 0 and ksize == 3:
        return np.array([0.25, 0.5, 0.25], dtype=F32)
    if sigma <= 0:
        sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8
    x = np.arange(ksize, dtype=F64) - (ksize - 1) * 0.5
    t = np.exp(-0.5 / (sigma * sigma) * x * x)
    return (t / t.sum()).astype(F32)


def gaussian_blur(img: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """Separable f32 blur, rows then columns, BORDER_REFLECT_101."""
    kern = gaussian_kernel(ksize, sigma)
    r = ksize // 2
    img = np.asarray(img, dtype=F32)
    H, W = img.shape
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    acc = np.zeros((H, W), dtype=F32)
    for i in range(ksize):
        acc = acc + kern[i] * p[:, i:i + W]
    p = np.pad(acc, ((r, r), (0, 0)), mode="reflect")
    out = np.zeros((H, W), dtype=F32)
    for i in range(ksize):
        out = out + kern[i] * p[i:i + H, :]
    return out


def _resize_taps(S: int, D: int, coef_dtype=F64):
    """INTER_LINEAR source index / fraction per destination index."""
    d = np.arange(D, dtype=F64)
    src = (d + 0.5) * (float(S) / D) - 0.5
    if coef_dtype == F32:
        src = src.astype(F32)
    s = np.floor(src).astype(np.int64)
    f = (src - s).astype(coef_dtype)
    lo = s < 0
    s[lo] = 0
    f[lo] = 0
    hi = s >= S - 1
    s[hi] = S - 1
    f[hi] = 0
    return s, f


def resize_linear(img: np.ndarray, h: int, w: int, coef_dtype=F64) -> np.ndarray:
    """cv::resize(INTER_LINEAR) on an (H,W) or (H,W,C) f32 array: horizontal
    pass then vertical.  coef_dtype=F64 mirrors the wheel's default (IPP) path,
    F32 mirrors OpenCV's own code."""
    img = np.asarray(img, dtype=F32)
    H, W = img.shape[:2]
    sx, fx = _resize_taps(W, w, coef_dtype)
    sy, fy = _resize_taps(H, h, coef_dtype)
    sx1 = np.minimum(sx + 1, W - 1)
    sy1 = np.minimum(sy + 1, H - 1)
    if img.ndim == 3:
        fxb = fx[None, :, None]
        fyb = fy[:, None, None]
    else:
        fxb = fx[None, :]
        fyb = fy[:, None]
    a = img[:, sx].astype(coef_dtype)
    b = img[:, sx1].astype(coef_dtype)
    hor = ((1 - fxb) * a + fxb * b).astype(F32)
    a = hor[sy].astype(coef_dtype)
    b = hor[sy1].astype(coef_dtype)
    return ((1 - fyb) * a + fyb * b).astype(F32)


def pyramid_image(img: np.ndarray, layer: dict, coef_dtype=F64) -> np.ndarray:
    blurred = gaussian_blur(np.asarray(img, dtype=F32), layer["ksize"], layer["sigma"])
    return resize_linear(blurred, layer["h"], layer["w"], coef_dtype)


# ----------------------------------------------------------------------------
# F4: polynomial expansion
# ----------------------------------------------------------------------------
def poly_exp_setup(n: int, sigma: float):
    """Returns g, xg, xxg (f32, index 0..n for offsets 0..n) and ig11, ig03, ig33, ig55."""
    if sigma < FLT_EPSILON:
        sigma = n * 0.3
    xs = np.arange(-n, n + 1)
    g = np.exp(-(xs * xs) / (2.0 * sigma * sigma)).astype(F32)
    s = 1.0 / g.astype(F64).sum()
    g = (g.astype(F64) * s).astype(F32)
    xg = (xs * g.astype(F64)).astype(F32)
    xxg = (xs * xs * g.astype(F64)).astype(F32)
    G = np.zeros((6, 6), dtype=F64)
    for iy, y in enumerate(xs):
        for ix, x in enumerate(xs):
            gg = F32(g[iy] * g[ix])
            G[0, 0] += gg
            G[1, 1] += F32(F32(gg * F32(x)) * F32(x))
            G[3, 3] += F32(F32(F32(F32(gg * F32(x)) * F32(x)) * F32(x)) * F32(x))
            G[5, 5] += F32(F32(F32(F32(gg * F32(x)) * F32(x)) * F32(y)) * F32(y))
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    inv = np.linalg.inv(G)
    return (g[n:].copy(), xg[n:].copy(), xxg[n:].copy(),
            float(inv[1, 1]), float(inv[0, 3]), float(inv[3, 3]), float(inv[5, 5]))


def poly_exp(I: np.ndarray, n: int, sigma: float) -> np.ndarray:
    """FarnebackPolyExp: (h,w) f32 -> (h,w,5) f32."""
    I = np.asarray(I, dtype=F32)
    h, w = I.shape
    g, xg, xxg, ig11, ig03, ig33, ig55 = poly_exp_setup(n, sigma)
    ys = np.arange(h)
    # vertical pass, f32, rows replicate
    r0 = I * g[0]
    r1 = np.zeros_like(I)
    r2 = np.zeros_like(I)
    for k in range(1, n + 1):
        s0 = I[np.maximum(ys - k, 0)]
        s1 = I[np.minimum(ys + k, h - 1)]
        p = s0 + s1
        r0 = r0 + g[k] * p
        r1 = r1 + xg[k] * (s1 - s0)
        r2 = r2 + xxg[k] * p
    # horizontal pass, double accumulators, columns replicate
    xsr = np.arange(w)
    b1 = (r0 * g[0]).astype(F64)
    b3 = (r1 * g[0]).astype(F64)
    b5 = (r2 * g[0]).astype(F64)
    b2 = np.zeros((h, w), dtype=F64)
    b4 = np.zeros((h, w), dtype=F64)
    b6 = np.zeros((h, w), dtype=F64)
    for k in range(1, n + 1):
        xp = np.minimum(xsr + k, w - 1)
        xm = np.maximum(xsr - k, 0)
        tg = (r0[:, xp] + r0[:, xm]).astype(F64)
        b1 += tg * F64(g[k])
        b4 += tg * F64(xxg[k])
        b2 += ((r0[:, xp] - r0[:, xm]) * xg[k]).astype(F64)
        b3 += ((r1[:, xp] + r1[:, xm]) * g[k]).astype(F64)
        b6 += ((r1[:, xp] - r1[:, xm]) * xg[k]).astype(F64)
        b5 += ((r2[:, xp] + r2[:, xm]) * g[k]).astype(F64)
    R = np.empty((h, w, 5), dtype=F32)
    R[..., 0] = (b3 * ig11).astype(F32)
    R[..., 1] = (b2 * ig11).astype(F32)
    R[..., 2] = (b1 * ig03 + b5 * ig33).astype(F32)
    R[..., 3] = (b1 * ig03 + b4 * ig33).astype(F32)
    R[..., 4] = (b6 * ig55).astype(F32)
    return R


# ----------------------------------------------------------------------------
# F5: update matrices
# ----------------------------------------------------------------------------
def border_scale(h: int, w: int) -> np.ndarray:
    sx = np.ones(w, dtype=F32)
    sy = np.ones(h, dtype=F32)
    for i in range(min(5, w)):
        sx[i] *= BORDER[i]
        sx[w - 1 - i] *= BORDER[i]
    for i in range(min(5, h)):
        sy[i] *= BORDER[i]
        sy[h - 1 - i] *= BORDER[i]
    # OpenCV multiplies in the order (x<B)*(x>=w-B)*(y<B)*(y>=h-B)
    return (sx[None, :] * sy[:, None]).astype(F32)


def update_matrices(R0: np.ndarray, R1: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """FarnebackUpdateMatrices: R0,R1 (h,w,5), flow (h,w,2) -> M (h,w,5), all f32."""
    h, w = flow.shape[:2]
    xs = np.arange(w, dtype=F32)[None, :]
    ys = np.arange(h, dtype=F32)[:, None]
    dx = flow[..., 0].astype(F32)
    dy = flow[..., 1].astype(F32)
    fx = (xs + dx).astype(F32)
    fy = (ys + dy).astype(F32)
    x1f = np.floor(fx)
    y1f = np.floor(fy)
    # guard int conversion against inf/nan/huge
    x1 = np.clip(np.nan_to_num(x1f, nan=-1e9), -2**30, 2**30).astype(np.int64)
    y1 = np.clip(np.nan_to_num(y1f, nan=-1e9), -2**30, 2**30).astype(np.int64)
    fx = (fx - x1f).astype(F32)
    fy = (fy - y1f).astype(F32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    x1c = np.clip(x1, 0, max(w - 2, 0))
    y1c = np.clip(y1, 0, max(h - 2, 0))
    one = F32(1)
    a00 = ((one - fx) * (one - fy)).astype(F32)
    a01 = (fx * (one - fy)).astype(F32)
    a10 = ((one - fx) * fy).astype(F32)
    a11 = (fx * fy).astype(F32)
    x2c = np.minimum(x1c + 1, w - 1)
    y2c = np.minimum(y1c + 1, h - 1)

    def samp(c):
        p = R1[..., c]
        return (a00 * p[y1c, x1c] + a01 * p[y1c, x2c]
                + a10 * p[y2c, x1c] + a11 * p[y2c, x2c]).astype(F32)

    half = F32(0.5)
    quarter = F32(0.25)
    r2 = np.where(inside, samp(0), F32(0)).astype(F32)
    r3 = np.where(inside, samp(1), F32(0)).astype(F32)
    r4 = np.where(inside, (R0[..., 2] + samp(2)) * half, R0[..., 2]).astype(F32)
    r5 = np.where(inside, (R0[..., 3] + samp(3)) * half, R0[..., 3]).astype(F32)
    r6 = np.where(inside, (R0[..., 4] + samp(4)) * quarter, R0[..., 4] * half).astype(F32)
    r2 = ((R0[..., 0] - r2) * half).astype(F32)
    r3 = ((R0[..., 1] - r3) * half).astype(F32)
    r2 = (r2 + r4 * dy + r6 * dx).astype(F32)
    r3 = (r3 + r6 * dy + r5 * dx).astype(F32)
    sc = border_scale(h, w)
    r2 = r2 * sc
    r3 = r3 * sc
    r4 = r4 * sc
    r5 = r5 * sc
    r6 = r6 * sc
    M = np.empty((h, w, 5), dtype=F32)
    M[..., 0] = r4 * r4 + r6 * r6
    M[..., 1] = (r4 + r5) * r6
    M[..., 2] = r5 * r5 + r6 * r6
    M[..., 3] = r4 * r2 + r6 * r3
    M[..., 4] = r6 * r2 + r5 * r3
    return M


# ----------------------------------------------------------------------------
# F6: box blur of M + 2x2 solve
# ----------------------------------------------------------------------------
def box_blur5(M: np.ndarray, winsize: int) -> np.ndarray:
    """(2m+1)^2 replicate-border window SUM in double times 1/winsize^2."""
    m = winsize // 2
    h, w = M.shape[:2]
    Md = M.astype(F64)
    ys = np.arange(h)
    xs = np.arange(w)
    v = np.zeros_like(Md)
    for d in range(-m, m + 1):
        v += Md[np.clip(ys + d, 0, h - 1)]
    s = np.zeros_like(Md)
    for d in range(-m, m + 1):
        s += v[:, np.clip(xs + d, 0, w - 1)]
    return s * (1.0 / (winsize * winsize))


def blur_solve(M: np.ndarray, winsize: int) -> np.ndarray:
    """FarnebackUpdateFlow_Blur: M (h,w,5) -> flow (h,w,2) f32."""
    B = box_blur5(M, winsize)
    g11, g12, g22, h1, h2 = (B[..., i] for i in range(5))
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    flow = np.empty(M.shape[:2] + (2,), dtype=F32)
    flow[..., 0] = ((g11 * h2 - g12 * h1) * idet).astype(F32)
    flow[..., 1] = ((g22 * h1 - g12 * h2) * idet).astype(F32)
    return flow


# ----------------------------------------------------------------------------
# F3 / F7: the layer loop
# ----------------------------------------------------------------------------
def calc_optical_flow_farneback(prev: np.ndarray, nxt: np.ndarray, pyr_scale=0.3, levels=5,
                                winsize=15, iterations=5, poly_n=5, poly_sigma=5.0,
                                flags=0, coef_dtype=F64, trace=None) -> np.ndarray:
    """Returns flow (H,W,2) f32.  ``trace``: optional list that receives one
    dict of intermediates per layer (for stage-by-stage kernel diffs)."""
    if flags != 0:
        raise ValueError("only flags=0 (box window, no initial flow) is on the reference path")
    prev = np.asarray(prev)
    nxt = np.asarray(nxt)
    H, W = prev.shape
    flow = None
    for layer in level_plan(H, W, pyr_scale, levels):
        h, w = layer["h"], layer["w"]
        if flow is None:
            flow = np.zeros((h, w, 2), dtype=F32)
        else:
            flow = resize_linear(flow, h, w, coef_dtype)
            flow = (flow.astype(F64) * (1.0 / pyr_scale)).astype(F32)
        I0 = pyramid_image(prev, layer, coef_dtype)
        I1 = pyramid_image(nxt, layer, coef_dtype)
        R0 = poly_exp(I0, poly_n, poly_sigma)
        R1 = poly_exp(I1, poly_n, poly_sigma)
        rec = dict(layer=layer, I0=I0, I1=I1, R0=R0, R1=R1, flow_init=flow.copy())
        M = update_matrices(R0, R1, flow)
        rec["M0"] = M
        for i in range(iterations):
            flow = blur_solve(M, winsize)
            if i < iterations - 1:
                M = update_matrices(R0, R1, flow)
        rec["flow"] = flow.copy()
        if trace is not None:
            trace.append(rec)
    return flow
