"""Seeded synthetic inputs shared by the tests, ``bench.py`` and ``smoke()``.

The reference ships no data (its config points at the author's disk,
/root/reference/Optical_flow/config.yaml:1-2, 28), so sweeps are modelled on
the CARLA sensor it records with (/root/reference/single_target_simultion.py:
63-73: ray-cast LiDAR, 100 m range, vertical FOV [-30, +15] deg, mounted
2.5 m above the vehicle origin) and BEV pairs on SURVEY.md §8(d).
numpy only; every function is a pure function of its integer seeds.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------
# BEV frame pairs (cfg3 / cfg4: pure flow throughput)
# ----------------------------------------------------------------------------
def bev_pair(pair_idx: int, H: int = 1024, W: int = 1024, max_disp: int = 3):
    """Two uint8 (H,W) frames: 20-200 rectangles of intensity U[60,255] over
    zeros, each displaced by an integer U[-max_disp, max_disp] px in frame 2."""
    rng = np.random.default_rng(pair_idx)
    a = np.zeros((H, W), dtype=np.uint8)
    b = np.zeros((H, W), dtype=np.uint8)
    n = int(rng.integers(20, 201))
    lim = max(4, min(H, W) // 16)
    for _ in range(n):
        h = int(rng.integers(3, lim))
        w = int(rng.integers(3, lim))
        y = int(rng.integers(max_disp, H - h - max_disp))
        x = int(rng.integers(max_disp, W - w - max_disp))
        v = int(rng.integers(60, 256))
        dy, dx = (int(t) for t in rng.integers(-max_disp, max_disp + 1, 2))
        a[y:y + h, x:x + w] = v
        b[y + dy:y + dy + h, x + dx:x + dx + w] = v
    return a, b


def bev_pair_sparse(pair_idx: int, H: int = 1024, W: int = 1024, max_disp: int = 3):
    """A mover-realistic pair: 4-10 vehicle-sized rectangles (10-40 px) over zeros, each displaced by an
    integer U[-max_disp, max_disp] px — about 10^4 moving cells per 1024x1024 pair (~10^3
    cells per mover) instead of the 1.5-2.6 x 10^5 of the dense throughput frames."""
    rng = np.random.default_rng(1_000_003 + pair_idx)
    a = np.zeros((H, W), dtype=np.uint8)
    b = np.zeros((H, W), dtype=np.uint8)
    for _ in range(int(rng.integers(4, 11))):
        h = int(rng.integers(10, 41))
        w = int(rng.integers(10, 41))
        y = int(rng.integers(max_disp, H - h - max_disp))
        x = int(rng.integers(max_disp, W - w - max_disp))
        v = int(rng.integers(60, 256))
        dy, dx = (int(t) for t in rng.integers(-max_disp, max_disp + 1, 2))
        a[y:y + h, x:x + w] = v
        b[y + dy:y + dy + h, x + dx:x + dx + w] = v
    return a, b


def bev_pairs(start: int, count: int, H: int = 1024, W: int = 1024, sparse: bool = False):
    """(prev uint8[count,H,W], next uint8[count,H,W]) for pairs start..start+count-1."""
    prev = np.empty((count, H, W), dtype=np.uint8)
    nxt = np.empty((count, H, W), dtype=np.uint8)
    for i in range(count):
        prev[i], nxt[i] = (bev_pair_sparse if sparse else bev_pair)(start + i, H, W)
    return prev, nxt


def textured_pair(seed: int, H: int, W: int, shift=(2, -3)):
    """Smooth random texture translated by an integer shift (well-conditioned flow)."""
    rng = np.random.default_rng(seed)
    pad = 16
    base = rng.uniform(0, 255, (H + 2 * pad, W + 2 * pad))
    for _ in range(3):          # cheap separable smoothing, no cv2 dependency
        base = (np.roll(base, 1, 0) + base + np.roll(base, -1, 0)) / 3
        base = (np.roll(base, 1, 1) + base + np.roll(base, -1, 1)) / 3
        base = (np.roll(base, 2, 0) + base + np.roll(base, -2, 0)) / 3
        base = (np.roll(base, 2, 1) + base + np.roll(base, -2, 1)) / 3
    base = (base - base.min()) / (base.max() - base.min()) * 255
    a = base[pad:pad + H, pad:pad + W]
    b = base[pad - shift[0]:pad - shift[0] + H, pad - shift[1]:pad - shift[1] + W]
    return a.astype(np.uint8), b.astype(np.uint8)


# ----------------------------------------------------------------------------
# LiDAR sweeps (cfg1 / cfg2 / cfg5)
# ----------------------------------------------------------------------------
def _ray_boxes(dirs: np.ndarray, boxes: np.ndarray, t_best: np.ndarray):
    """Slab test of rays from the origin against axis-aligned boxes (nb,6:
    xmin,xmax,ymin,ymax,zmin,zmax); updates t_best in place."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
    for bx in boxes:
        lo = np.array([bx[0], bx[2], bx[4]])
        hi = np.array([bx[1], bx[3], bx[5]])
        t0 = lo * inv
        t1 = hi * inv
        tmin = np.minimum(t0, t1).max(axis=1)
        tmax = np.maximum(t0, t1).min(axis=1)
        hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.05)
        np.minimum(t_best, np.where(hit, tmin, np.inf), out=t_best)


def scene(seq: int, n_movers: int = 1, n_static: int = 12, extent: float = 45.0):
    """Static boxes and movers (start position, velocity m/s) of sequence ``seq``."""
    rng = np.random.default_rng(7919 * (seq + 1))
    statics = []
    for _ in range(n_static):
        cx, cy = rng.uniform(-extent, extent, 2)
        if abs(cx) < 6 and abs(cy) < 6:
            cx += 12
        sx, sy = rng.uniform(3, 10, 2)
        hgt = rng.uniform(3, 10)
        statics.append([cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2, -2.5, -2.5 + hgt])
    movers = []
    for _ in range(n_movers):
        cx, cy = rng.uniform(-0.6 * extent, 0.6 * extent, 2)
        if abs(cx) < 5 and abs(cy) < 5:
            cy += 10
        speed = rng.uniform(0.5, 15.0)
        ang = rng.uniform(0, 2 * np.pi)
        movers.append(dict(p0=np.array([cx, cy]), v=speed * np.array([np.cos(ang), np.sin(ang)]),
                           size=np.array([4.5, 1.8, 1.5])))
    return np.array(statics), movers


def lidar_sweep(seq: int, frame: int, beams: int = 32, n_points: int = 60_000, n_movers: int = 1,
                dt: float = 0.1, max_range: float = 100.0, extent: float = 45.0) -> np.ndarray:
    """One sweep as float32 (N,4) x,y,z,intensity in the sensor frame (the layout
    CARLA emits, single_target_simultion.py:260).  Ground is the plane z=-2.5
    with N(0, 0.02) roughness; returns beyond ``max_range`` are dropped, so N is
    a little under ``n_points``."""
    statics, movers = scene(seq, n_movers, extent=extent)
    rng = np.random.default_rng(frame + 1000 * seq)
    n_az = max(8, int(round(n_points * 1.35 / beams)))
    elev = np.deg2rad(np.linspace(-30.0, 15.0, beams))
    az = np.linspace(0.0, 2 * np.pi, n_az, endpoint=False) + rng.uniform(0, 2 * np.pi / n_az)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    dirs = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :],
                     np.broadcast_to(se, (beams, n_az))], axis=-1).reshape(-1, 3)
    t = np.full(len(dirs), np.inf)
    down = dirs[:, 2] < -1e-6
    tg = np.where(down, -2.5 / np.where(down, dirs[:, 2], -1.0), np.inf)
    np.minimum(t, tg, out=t)
    boxes = [statics]
    mv = []
    for m in movers:
        c = m["p0"] + m["v"] * (frame * dt)
        sx, sy, sz = m["size"]
        mv.append([c[0] - sx / 2, c[0] + sx / 2, c[1] - sy / 2, c[1] + sy / 2, -2.5, -2.5 + sz])
    if mv:
        boxes.append(np.array(mv))
    _ray_boxes(dirs, np.concatenate(boxes), t)
    keep = np.isfinite(t) & (t < max_range)
    pts = dirs[keep] * t[keep, None]
    pts += rng.normal(0.0, 0.02, pts.shape)
    out = np.empty((len(pts), 4), dtype=np.float32)
    out[:, :3] = pts
    out[:, 3] = rng.uniform(0.1, 1.0, len(pts))
    return out


SWEEP_CONFIGS = {
    # BASELINE.json configs[0], [1], [4]
    "cfg1": dict(beams=32, n_points=60_000, n_movers=1, grid_resolution=(0.25, 0.25),
                 x_range=(-50.0, 50.0), y_range=(-50.0, 50.0)),
    "cfg2": dict(beams=64, n_points=120_000, n_movers=10, grid_resolution=(0.125, 0.125),
                 x_range=(-50.0, 50.0), y_range=(-50.0, 50.0)),
    "cfg5": dict(beams=128, n_points=240_000, n_movers=10, grid_resolution=(0.1, 0.1),
                 x_range=(-51.2, 51.2), y_range=(-51.2, 51.2)),
}
