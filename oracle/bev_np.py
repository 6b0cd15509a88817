"""numpy restatement of the reference's per-frame preprocessing — TEST ORACLE.

Follows /root/reference/Optical_flow/main.py:
  filter_points_in_roi      main.py:30-36
  increase_point_density    main.py:38-57   (noise is supplied by the caller so both
                                             paths see the same expanded array)
  compute_bev_grid          main.py:98-126

``compute_bev_grid_loops`` is the loop-for-loop restatement (small cases);
``compute_bev_grid`` is the vectorised one used at real sizes.  Both are pinned
bit-exactly (uint8) against the reference function itself, imported in the
build container, through tests/golden/ (see tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def filter_points_in_roi(points: np.ndarray, roi_bounds) -> np.ndarray:
    """main.py:30-36 — closed intervals on all six bounds."""
    x_min, x_max, y_min, y_max, z_min, z_max = roi_bounds
    p = points
    keep = ((p[:, 0] >= x_min) & (p[:, 0] <= x_max) &
            (p[:, 1] >= y_min) & (p[:, 1] <= y_max) &
            (p[:, 2] >= z_min) & (p[:, 2] <= z_max))
    return p[keep]


def increase_point_density(points: np.ndarray, expansion_factor: int = 2, noise_std: float = 0.01,
                           noise: np.ndarray | None = None, rng: np.random.Generator | None = None):
    """main.py:38-57 — each point repeated ``expansion_factor`` times (consecutive
    copies) plus N(0, noise_std) on x, y and z.  The reference draws from the
    unseeded global RNG; here the harness owns the noise."""
    rep = np.repeat(np.asarray(points, dtype=np.float64), expansion_factor, axis=0)
    if noise is None:
        rng = rng or np.random.default_rng(0)
        noise = rng.normal(scale=noise_std, size=rep.shape)
    return rep + noise


def cast_u8(v: np.ndarray) -> np.ndarray:
    """numpy's float64 -> uint8 cast on x86-64 as the reference relies on it at
    main.py:123: truncate toward zero through int32 (out of range / NaN ->
    INT_MIN), keep the low byte."""
    v = np.asarray(v, dtype=np.float64)
    ok = np.isfinite(v) & (np.abs(v) < 2147483648.0)
    t = np.where(ok, np.trunc(np.where(ok, v, 0.0)), -2147483648.0).astype(np.int64)
    return (t & 0xFF).astype(np.uint8)


def _bins(lo, hi, step):
    return len(np.arange(lo, hi, step))


def compute_bev_grid_loops(points, grid_resolution, x_range, y_range, a=0.5, b=0.5, h_max=5.0):
    """Loop-for-loop restatement of main.py:98-126 (slow; small cases only)."""
    w, h = grid_resolution
    nx = _bins(x_range[0], x_range[1], w)
    ny = _bins(y_range[0], y_range[1], h)
    cells = [[[] for _ in range(ny)] for _ in range(nx)]
    for x, y, z in points:
        xi = int((x - x_range[0]) / w)          # truncation toward zero
        yi = int((y - y_range[0]) / h)
        if 0 <= xi < nx and 0 <= yi < ny:
            cells[xi][yi].append(z)
    vals = np.zeros((nx, ny))
    for i in range(nx):
        for j in range(ny):
            hs = np.array(cells[i][j])
            if len(hs) > 0:
                vals[i, j] = (a * np.mean(hs) + b * np.std(hs)) / h_max
    with np.errstate(all="ignore"):
        vals = vals / vals.max()
        return cast_u8(vals * 255)


def cell_indices(points, grid_resolution, x_range, y_range):
    """(xi, yi, keep) with the reference's truncation-toward-zero binning
    (main.py:106-108): x in (lo-w, lo) lands in cell 0."""
    w, h = grid_resolution
    nx = _bins(x_range[0], x_range[1], w)
    ny = _bins(y_range[0], y_range[1], h)
    p = np.asarray(points, dtype=np.float64)
    with np.errstate(all="ignore"):
        qx = np.trunc((p[:, 0] - x_range[0]) / w)
        qy = np.trunc((p[:, 1] - y_range[0]) / h)
    keep = (qx >= 0) & (qx < nx) & (qy >= 0) & (qy < ny)   # NaN fails both
    xi = np.where(keep, qx, 0).astype(np.int64)
    yi = np.where(keep, qy, 0).astype(np.int64)
    return xi, yi, keep, nx, ny


def bev_cell_values(points, grid_resolution, x_range, y_range, a=0.5, b=0.5, h_max=5.0):
    """Per-cell (a*mean + b*std)/h_max before normalisation, float64 (nx, ny)."""
    xi, yi, keep, nx, ny = cell_indices(points, grid_resolution, x_range, y_range)
    z = np.asarray(points, dtype=np.float64)[:, 2][keep]
    cell = (xi * ny + yi)[keep]
    cnt = np.bincount(cell, minlength=nx * ny).astype(np.float64)
    s = np.bincount(cell, weights=z, minlength=nx * ny)
    with np.errstate(all="ignore"):
        mean = np.where(cnt > 0, s / cnt, 0.0)
        dev = z - mean[cell]
        var = np.bincount(cell, weights=dev * dev, minlength=nx * ny) / np.where(cnt > 0, cnt, 1.0)
    vals = np.where(cnt > 0, (a * mean + b * np.sqrt(var)) / h_max, 0.0)
    return vals.reshape(nx, ny), cnt.reshape(nx, ny)


def compute_bev_grid(points, grid_resolution, x_range, y_range, a=0.5, b=0.5, h_max=5.0):
    """Vectorised restatement of main.py:98-126 -> uint8 (nx, ny), axis 0 = x."""
    vals, _ = bev_cell_values(points, grid_resolution, x_range, y_range, a, b, h_max)
    with np.errstate(all="ignore"):
        vals = vals / vals.max()
        return cast_u8(vals * 255)
