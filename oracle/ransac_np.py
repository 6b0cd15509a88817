"""Restatement of Open3D ``PointCloud.segment_plane`` — TEST ORACLE, PARITY UNPINNED.

The reference calls ``segment_plane(distance_threshold=0.5, ransac_n=5,
num_iterations=5000)`` at /root/reference/Optical_flow/main.py:73 and drops the
returned inliers (main.py:74-75).  open3d is unpinned by the reference
(README.md:26), absent from this image and not vendored, and the algorithm is
randomised, so nothing here can be checked against the real library: this file
restates the published algorithm (geometry/PointCloudSegmentation.cpp:
GetPlaneFromPoints, EvaluateRANSACBasedOnDistance, SegmentPlane) and fixes the
one thing Open3D leaves to its RNG — which points each hypothesis samples — as
a counter-based integer hash so that the CUDA path and this oracle evaluate the
SAME hypotheses and can be compared exactly:

  * sample j of iteration i = mix64(seed, i*ATTEMPTS + t) mod N, taking
    attempts t = 0,1,.. until ``ransac_n`` distinct indices are found
    (at most ATTEMPTS tries, else the hypothesis is skipped);
  * plane through the samples: 3 points -> triangle normal; more -> centroid +
    covariance cofactor normal (GetPlaneFromPoints);
  * score: inlier <=> |a x + b y + c z + d| < threshold, fp64, evaluated as
    ((a*x + b*y) + c*z) + d with every operation rounded (no FMA);
    error = sum of inlier distances; rmse = error / sqrt(count);
  * best = most inliers, ties by lower rmse, then by lower iteration index
    (Open3D's order under OpenMP is unspecified);
  * all ``num_iterations`` hypotheses are scored (Open3D's probabilistic early
    exit only shortens the search; it never changes how a hypothesis is scored);
  * final inliers are those of the best hypothesis plane; the returned plane is
    refit on them (GetPlaneFromPoints).
"""
from __future__ import annotations

import numpy as np

ATTEMPTS = 16
U64 = np.uint64


def mix64(seed, ctr):
    """splitmix64 finaliser of (seed + ctr * golden), vectorised, uint64."""
    with np.errstate(over="ignore"):
        z = (np.asarray(seed, dtype=U64) + (np.asarray(ctr, dtype=U64) + U64(1)) * U64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


def sample_indices(seed: int, num_iterations: int, ransac_n: int, n_points: int):
    """(idx int64[num_iterations, ransac_n], ok bool[num_iterations])."""
    idx = -np.ones((num_iterations, ransac_n), dtype=np.int64)
    filled = np.zeros(num_iterations, dtype=np.int64)
    it = np.arange(num_iterations, dtype=np.uint64)
    for t in range(ATTEMPTS):
        cand = (mix64(seed, it * U64(ATTEMPTS) + U64(t)) % U64(n_points)).astype(np.int64)
        dup = (idx == cand[:, None]).any(axis=1)
        take = (~dup) & (filled < ransac_n)
        rows = np.nonzero(take)[0]
        idx[rows, filled[rows]] = cand[rows]
        filled[rows] += 1
    return idx, filled >= ransac_n


def plane_from_points(P: np.ndarray) -> np.ndarray:
    """GetPlaneFromPoints on (..., m, 3) fp64 -> (..., 4); zero plane if degenerate."""
    P = np.asarray(P, dtype=np.float64)
    m = P.shape[-2]
    if m == 3:
        e0 = P[..., 1, :] - P[..., 0, :]
        e1 = P[..., 2, :] - P[..., 0, :]
        abc = np.cross(e0, e1)
        c = P[..., 0, :]
    else:
        c = P.sum(axis=-2) / m
        r = P - c[..., None, :]
        xx = (r[..., 0] * r[..., 0]).sum(-1)
        xy = (r[..., 0] * r[..., 1]).sum(-1)
        xz = (r[..., 0] * r[..., 2]).sum(-1)
        yy = (r[..., 1] * r[..., 1]).sum(-1)
        yz = (r[..., 1] * r[..., 2]).sum(-1)
        zz = (r[..., 2] * r[..., 2]).sum(-1)
        det_x = yy * zz - yz * yz
        det_y = xx * zz - xz * xz
        det_z = xx * yy - xy * xy
        ax = np.stack([det_x, xz * yz - xy * zz, xy * yz - xz * yy], -1)
        ay = np.stack([xz * yz - xy * zz, det_y, xy * xz - yz * xx], -1)
        az = np.stack([xy * yz - xz * yy, xy * xz - yz * xx, det_z], -1)
        use_x = (det_x > det_y) & (det_x > det_z)
        use_y = (~use_x) & (det_y > det_z)
        abc = np.where(use_x[..., None], ax, np.where(use_y[..., None], ay, az))
    norm = np.sqrt((abc * abc).sum(-1))
    with np.errstate(all="ignore"):
        abc_n = abc / norm[..., None]
    d = -(abc_n * c).sum(-1)
    plane = np.concatenate([abc_n, d[..., None]], -1)
    bad = ~(norm > 0) | ~np.isfinite(plane).all(-1)
    plane[bad] = 0.0
    return plane


def point_plane_distance(points: np.ndarray, plane: np.ndarray) -> np.ndarray:
    """|((a*x + b*y) + c*z) + d| with every op rounded, fp64."""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    return np.abs(((plane[0] * x + plane[1] * y) + plane[2] * z) + plane[3])


def score_planes(points: np.ndarray, planes: np.ndarray, threshold: float):
    """(count int64[H], err float64[H]) for every hypothesis."""
    pts = np.asarray(points, dtype=np.float64)
    cnt = np.zeros(len(planes), dtype=np.int64)
    err = np.zeros(len(planes), dtype=np.float64)
    for h, pl in enumerate(planes):
        if not pl.any():
            continue
        d = point_plane_distance(pts, pl)
        m = d < threshold
        cnt[h] = int(m.sum())
        err[h] = float(d[m].sum())
    return cnt, err


def select_best(cnt: np.ndarray, err: np.ndarray) -> int:
    """Most inliers; ties by lower rmse = err/sqrt(cnt); then lower index.  -1 if none."""
    with np.errstate(all="ignore"):
        rmse = np.where(cnt > 0, err / np.sqrt(np.maximum(cnt, 1)), np.inf)
    best = -1
    for h in range(len(cnt)):
        if cnt[h] <= 0:
            continue
        if best < 0 or cnt[h] > cnt[best] or (cnt[h] == cnt[best] and rmse[h] < rmse[best]):
            best = h
    return best


def segment_plane(points: np.ndarray, distance_threshold=0.5, ransac_n=5, num_iterations=5000, seed=0):
    """-> (plane[4] refit on the inliers, inlier_mask bool[N], best hypothesis plane[4])."""
    pts = np.asarray(points, dtype=np.float64)[:, :3]
    n = len(pts)
    if n < ransac_n:
        raise ValueError("There must be at least 'ransac_n' points.")
    idx, ok = sample_indices(seed, num_iterations, ransac_n, n)
    planes = plane_from_points(pts[np.where(ok[:, None], idx, 0)])
    planes[~ok] = 0.0
    cnt, err = score_planes(pts, planes, distance_threshold)
    best = select_best(cnt, err)
    if best < 0:
        return np.zeros(4), np.zeros(n, dtype=bool), np.zeros(4)
    mask = point_plane_distance(pts, planes[best]) < distance_threshold
    refit = plane_from_points(pts[mask][None])[0] if mask.sum() >= 3 else planes[best]
    return refit, mask, planes[best]
