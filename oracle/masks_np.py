"""numpy restatement of the velocity / mask stage — TEST ORACLE.

Follows /root/reference/Optical_flow/main.py:
  velocity scaling + curl     main.py:143-164  (inside compute_velocity_vectors)
  continuity_mask             main.py:224-228
  moving-cell filter (inline) main.py:596-609

``np.gradient`` is restated explicitly (central difference /2 in the interior,
one-sided first-order at the two edges) in the dtype numpy uses there: f32 in
f32 out.  Pinned against the reference functions through tests/golden/.
"""
from __future__ import annotations

import numpy as np


def gradient(f: np.ndarray, axis: int) -> np.ndarray:
    """np.gradient(f, axis=axis) with unit spacing, edge_order=1, same dtype."""
    f = np.asarray(f)
    f = np.moveaxis(f, axis, 0)
    out = np.empty_like(f)
    n = f.shape[0]
    if n < 2:
        raise ValueError("np.gradient needs at least 2 samples along the axis")
    out[1:-1] = (f[2:] - f[:-2]) / f.dtype.type(2.0)
    out[0] = f[1] - f[0]
    out[-1] = f[-1] - f[-2]
    return np.moveaxis(out, 0, axis)


def flow_to_velocity(flow: np.ndarray, x_range, y_range):
    """main.py:143-160: velocity = flow * pixel size (dt is ignored by the
    reference), plus the curl it returns as ``angular_velocity``."""
    H, W = flow.shape[:2]
    px = (x_range[1] - x_range[0]) / W          # main.py:147 uses shape[1] for x
    py = (y_range[1] - y_range[0]) / H
    vx = flow[..., 0] * px                       # python float * f32 array stays f32
    vy = flow[..., 1] * py
    ang = gradient(vy, 1) - gradient(vx, 0)
    return vx, vy, ang


def div_curl(vx: np.ndarray, vy: np.ndarray):
    div = gradient(vx, 1) + gradient(vy, 0)
    curl = gradient(vy, 1) - gradient(vx, 0)
    return div, curl


def continuity_mask(vx: np.ndarray, vy: np.ndarray, alpha_cont: float) -> np.ndarray:
    """main.py:224-228 -> int64 0/1."""
    div, curl = div_curl(vx, vy)
    return ((np.abs(div) <= alpha_cont) & (np.abs(curl) <= alpha_cont)).astype(np.int64)


def moving_cell_filter(vx: np.ndarray, vy: np.ndarray, mask: np.ndarray, thresh: float = 0.1):
    """main.py:600-609: filtered velocities (f64), magnitude, curl of the filtered
    field and the strict ``mag > 0.1`` valid mask."""
    vx_f = vx * mask                              # f32 * int64 -> f64
    vy_f = vy * mask
    mag = np.sqrt(vx_f ** 2 + vy_f ** 2)
    ang = gradient(vy_f, 1) - gradient(vx_f, 0)
    return vx_f, vy_f, mag, ang, mag > thresh


def _propagate(vx, vy, di, dj, alpha_p):
    """Forward scatter with last-writer-wins in row-major order (the reference's double loop,
    main.py:172-178 / 209-216), vectorised: the winner of a target cell is the LARGEST source
    rank that lands on it."""
    h, w = vx.shape
    i, j = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    # `i + np.floor(...)`: python int + np.float32 stays float32 (NEP 50); int() truncates
    with np.errstate(invalid="ignore", over="ignore"):
        ti = (i.astype(vx.dtype) + di)
        tj = (j.astype(vx.dtype) + dj)
    ok = np.isfinite(ti) & np.isfinite(tj)
    ti_i = np.where(ok, np.trunc(np.where(ok, ti, 0)), -1).astype(np.int64)
    tj_i = np.where(ok, np.trunc(np.where(ok, tj, 0)), -1).astype(np.int64)
    ok &= (ti_i >= 0) & (ti_i < h) & (tj_i >= 0) & (tj_i < w)
    rank = (i * w + j)[ok]
    target = (ti_i * w + tj_i)[ok]
    winner = np.full(h * w, -1, np.int64)
    np.maximum.at(winner, target, rank)
    pvx = np.zeros(h * w, vx.dtype)
    pvy = np.zeros(h * w, vy.dtype)
    hit = winner >= 0
    pvx[hit] = vx.ravel()[winner[hit]]
    pvy[hit] = vy.ravel()[winner[hit]]
    pvx, pvy = pvx.reshape(h, w), pvy.reshape(h, w)
    return ((np.abs(pvx - vx) <= alpha_p) & (np.abs(pvy - vy) <= alpha_p)).astype(np.int64)


def propagation_mask(vx, vy, dt, grid_resolution, alpha_p):
    """main.py:166-182 (defined by the reference, never called by its driver).  NaN / inf
    displacements make the reference's int() raise; here they simply do not propagate."""
    with np.errstate(invalid="ignore", over="ignore"):
        di = np.floor(vx * dt / grid_resolution[0])
        dj = np.floor(vy * dt / grid_resolution[1])
    return _propagate(vx, vy, di, dj, alpha_p)


def propagation_mask_with_acceleration(vx, vy, ax, ay, dt, grid_resolution, alpha_p):
    """main.py:184-221."""
    dx, dy = grid_resolution
    with np.errstate(invalid="ignore", over="ignore"):
        di = np.floor((vx * dt + 0.5 * ax * dt ** 2) / dx)
        dj = np.floor((vy * dt + 0.5 * ay * dt ** 2) / dy)
    return _propagate(vx, vy, di, dj, alpha_p)
