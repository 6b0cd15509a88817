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
