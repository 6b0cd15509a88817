"""Drop-in for the hot-path functions of the reference's ``Optical_flow/main.py``.

Same names, argument meaning, return types and error behaviour as the
reference, so its driver loop (process_multiple_frames, main.py:541-641) can
import these instead of its own:

    from datmo_using_optical_flow_b200.main import (
        filter_points_in_roi, increase_point_density, compute_bev_grid,
        preprocess_points, compute_velocity_vectors, continuity_mask,
        dbscan_clustering, extract_cluster_data)

numpy arrays in -> numpy arrays out (the reference's dtypes); CUDA tensors in
-> CUDA tensors out.  Every function runs on the GPU through libdatmo_b200;
there is no CPU fallback — without a B200 and the built library they raise.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import Engine, default_engine, farneback_params

__all__ = ["filter_points_in_roi", "increase_point_density", "compute_bev_grid", "preprocess_points",
           "preprocess_pcd", "compute_velocity_vectors", "continuity_mask", "propagation_mask",
           "propagation_mask_with_acceleration", "moving_cell_filter",
           "dbscan_clustering", "extract_cluster_data", "clusters_from_summary", "flow_to_clusters", "read_pcd"]


def _to_dev(eng: Engine, a, dtype=None):
    if isinstance(a, torch.Tensor):
        t = a.to(eng.tdev)
    else:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(eng.tdev, non_blocking=False)
    return t if dtype is None else t.to(dtype)


def _is_np(*xs) -> bool:
    return not any(isinstance(x, torch.Tensor) for x in xs)


# ---------------------------------------------------------------------------------------------
# preprocessing (main.py:30-126)
# ---------------------------------------------------------------------------------------------
def filter_points_in_roi(points, roi_bounds, engine=None):
    """main.py:30-36 — closed-interval box crop, order preserved (device compaction)."""
    eng = engine or default_engine()
    as_np = _is_np(points)
    pts = _to_dev(eng, points)
    if pts.dim() != 2 or pts.shape[1] < 3:
        raise ValueError("points must be (N, 3)")
    if not (pts.dtype == torch.float32 and pts.shape[1] == 4):
        pts = pts[:, :3].to(torch.float64)
    out = eng.roi_filter(pts, roi_bounds)
    return out.cpu().numpy() if as_np else out


def increase_point_density(points, expansion_factor=2, noise_std=0.01, noise=None, seed=None, engine=None):
    """main.py:38-57 — ``expansion_factor`` consecutive copies of every point plus
    N(0, noise_std).  ``noise`` (same shape as the result) makes the call reproducible;
    otherwise noise is drawn on the device (the reference uses numpy's unseeded global RNG)."""
    eng = engine or default_engine()
    as_np = _is_np(points)
    pts = _to_dev(eng, points)[:, :3].to(torch.float64)
    nz = None if noise is None else _to_dev(eng, noise, torch.float64)
    if seed is None:
        seed = int(np.random.randint(0, 2**31 - 1))
    out = eng.expand_points(pts, int(expansion_factor), noise_std, nz, seed)
    if as_np:
        eng.synchronize()
        return out.cpu().numpy()
    return out


def compute_bev_grid(points, grid_resolution, x_range, y_range, a=0.5, b=0.5, h_max=5.0, engine=None):
    """main.py:98-126 -> uint8 (nx, ny), axis 0 = x.  Bit-exact with the reference."""
    eng = engine or default_engine()
    as_np = _is_np(points)
    pts = _to_dev(eng, points)
    if pts.dim() != 2 or pts.shape[1] < 3:
        raise ValueError("points must be (N, 3)")
    if pts.dtype == torch.float32 and pts.shape[1] == 4:
        pass
    else:
        pts = pts[:, :3].to(torch.float64)
    bev = eng.bev_rasterize(pts, grid_resolution, x_range, y_range, a, b, h_max)
    if as_np:
        eng.synchronize()
        return bev.cpu().numpy()
    return bev


def preprocess_points(points_xyzw, grid_resolution, x_range, y_range, z_max, roi_bounds, seed=0, noise=None,
                      ground_mask=None, ransac=(0.5, 5, 5000), expansion_factor=10, noise_std=0.01, engine=None):
    """preprocess_pcd (main.py:59-95) from the loaded cloud on: flip x, RANSAC ground
    removal (main.py:73 hard-codes (0.5, 5, 5000)), ROI crop, x10 expansion, BEV.
    points_xyzw: float32 (N,4) as CARLA writes them.  Returns uint8 (nx,ny) or None
    when no point falls inside the ROI (main.py:84-86)."""
    eng = engine or default_engine()
    as_np = _is_np(points_xyzw)
    pts = _to_dev(eng, points_xyzw, torch.float32)
    if pts.dim() != 2 or pts.shape[1] not in (3, 4):
        raise ValueError("points must be (N, 3) or (N, 4)")
    if pts.shape[1] == 3:
        pts = torch.cat([pts, torch.zeros_like(pts[:, :1])], dim=1)
    nz = None if noise is None else _to_dev(eng, noise, torch.float64)
    gm = None if ground_mask is None else _to_dev(eng, ground_mask, torch.uint8)
    bev = eng.preprocess(pts, grid_resolution, x_range, y_range, z_max, roi_bounds, ransac[0], ransac[1], ransac[2],
                         seed, True, gm, expansion_factor, noise_std, nz)
    if bev is None:
        return None
    return bev.cpu().numpy() if as_np else bev


def read_pcd(pcd_file) -> np.ndarray:
    """Minimal .pcd reader (ascii / binary, x y z [+ anything]) replacing
    o3d.io.read_point_cloud at main.py:60 -> float32 (N,4)."""
    with open(pcd_file, "rb") as fh:
        fields, sizes, types, counts, npts, data_kind = [], [], [], [], 0, "ascii"
        while True:
            line = fh.readline()
            if not line:
                raise ValueError(f"{pcd_file}: truncated PCD header")
            tok = line.decode("ascii", "replace").strip().split()
            if not tok or tok[0].startswith("#"):
                continue
            key = tok[0].upper()
            if key == "FIELDS":
                fields = tok[1:]
            elif key == "SIZE":
                sizes = [int(t) for t in tok[1:]]
            elif key == "TYPE":
                types = tok[1:]
            elif key == "COUNT":
                counts = [int(t) for t in tok[1:]]
            elif key == "POINTS":
                npts = int(tok[1])
            elif key == "DATA":
                data_kind = tok[1].lower()
                break
        counts = counts or [1] * len(fields)
        if not all(f in fields for f in ("x", "y", "z")):
            raise ValueError(f"{pcd_file}: PCD has no x/y/z fields")
        if data_kind == "ascii":
            arr = np.loadtxt(fh, dtype=np.float64, ndmin=2)
            cols = np.cumsum([0] + counts)
            xyz = np.stack([arr[:, cols[fields.index(c)]] for c in "xyz"], axis=1)
        elif data_kind == "binary":
            code = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4",
                    ("I", 1): "i1", ("I", 2): "<i2", ("I", 4): "<i4"}
            dt = np.dtype([(f, code[(t.upper(), s)], (c,)) for f, s, t, c in zip(fields, sizes, types, counts)])
            rec = np.frombuffer(fh.read(npts * dt.itemsize), dtype=dt, count=npts)
            xyz = np.stack([rec[c][:, 0].astype(np.float64) for c in "xyz"], axis=1)
        else:
            raise ValueError(f"{pcd_file}: unsupported PCD DATA {data_kind}")
    out = np.zeros((len(xyz), 4), dtype=np.float32)
    out[:, :3] = xyz
    return out


def preprocess_pcd(pcd_file, grid_resolution, x_range, y_range, z_max, roi_bounds, engine=None, seed=0):
    """main.py:59-95 with the reference's signature."""
    bev = preprocess_points(read_pcd(pcd_file), grid_resolution, x_range, y_range, z_max, roi_bounds, seed=seed,
                            engine=engine)
    if bev is None:
        print(f"No ROI points for {pcd_file}. Adjust ROI bounds.")
        return None
    print(f"BEV grid computed for file: {pcd_file}")
    return bev


# ---------------------------------------------------------------------------------------------
# flow -> velocity -> masks (main.py:131-228, 596-609)
# ---------------------------------------------------------------------------------------------
def _bev_to_dev(eng, bev):
    t = _to_dev(eng, bev)
    if t.dtype not in (torch.uint8, torch.float32):
        t = t.to(torch.float32)
    return t


def compute_velocity_vectors(bev1, bev2, x_range, y_range, dt, farneback=None, engine=None):
    """main.py:131-164 -> (velocity_x, velocity_y, angular_velocity), float32 (H,W).
    ``dt`` is accepted and ignored, as in the reference.  ``farneback``: optional dict
    overriding the parameters hard-coded at main.py:132-140."""
    eng = engine or default_engine()
    as_np = _is_np(bev1, bev2)
    a, b = _bev_to_dev(eng, bev1), _bev_to_dev(eng, bev2)
    H, W = a.shape[-2:]
    params = farneback_params(**(farneback or {}))
    flow = eng.farneback(a, b, params)
    px = (x_range[1] - x_range[0]) / W      # main.py:147 divides the x range by shape[1]
    py = (y_range[1] - y_range[0]) / H
    vm = eng.velocity_mask(flow, px, py, 0.0, 0.1, want=("vx", "vy", "ang"))
    out = tuple(vm[k][0] if a.dim() == 2 else vm[k] for k in ("vx", "vy", "ang"))
    if as_np:
        eng.synchronize()
        return tuple(t.cpu().numpy() for t in out)
    return out


def _velocity_as_flow(eng, vx, vy):
    vx, vy = _to_dev(eng, vx, torch.float32), _to_dev(eng, vy, torch.float32)
    return torch.stack([vx, vy], dim=-1).contiguous()


def continuity_mask(vx, vy, alpha_cont, engine=None):
    """main.py:224-228 -> int64 0/1 (H,W)."""
    eng = engine or default_engine()
    as_np = _is_np(vx, vy)
    flow = _velocity_as_flow(eng, vx, vy)
    vm = eng.velocity_mask(flow, 1.0, 1.0, alpha_cont, 0.1, want=("mask",))
    m = vm["mask"][0] if flow.dim() == 3 else vm["mask"]
    if as_np:
        eng.synchronize()
        return m.cpu().numpy().astype(np.int64)
    return m.to(torch.int64)


def _propagation(vx, vy, ax, ay, dt, grid_resolution, alpha_p, engine):
    eng = engine or default_engine()
    as_np = _is_np(vx, vy)
    # the reference computes in the dtype of its inputs: f32 velocities stay f32, anything else is f64
    dtype = torch.float32 if (getattr(vx, "dtype", None) in (np.float32, torch.float32)) else torch.float64
    dev = [None if t is None else _to_dev(eng, t, dtype) for t in (vx, vy, ax, ay)]
    m = eng.propagation_mask(dev[0], dev[1], dt, grid_resolution, alpha_p, dev[2], dev[3])
    if as_np:
        eng.synchronize()
        return m.cpu().numpy().astype(np.int64)
    return m.to(torch.int64)


def propagation_mask(vx, vy, dt, grid_resolution, alpha_p, engine=None):
    """main.py:166-182 -> int64 0/1 (H,W): forward scatter of every cell's velocity (last source in
    row-major order wins), compared with the actual field."""
    return _propagation(vx, vy, None, None, dt, grid_resolution, alpha_p, engine)


def propagation_mask_with_acceleration(vx, vy, ax, ay, dt, grid_resolution, alpha_p, engine=None):
    """main.py:184-221: the same with the displacement (v dt + a dt^2 / 2)."""
    return _propagation(vx, vy, ax, ay, dt, grid_resolution, alpha_p, engine)


def moving_cell_filter(vx, vy, alpha_cont, thresh=0.1, engine=None):
    """The inline filter of process_multiple_frames (main.py:596-609) ->
    (vx_filtered f64, vy_filtered f64, velocity_magnitude f64, angular_velocity f64, valid_mask bool)."""
    eng = engine or default_engine()
    as_np = _is_np(vx, vy)
    flow = _velocity_as_flow(eng, vx, vy)
    vm = eng.velocity_mask(flow, 1.0, 1.0, alpha_cont, thresh, want=("vx_f", "vy_f", "valid"))
    sq = (lambda t: t[0]) if flow.dim() == 3 else (lambda t: t)
    g = eng.filtered_grids_f64(vm["vx_f"], vm["vy_f"])       # f64 like the reference's arrays, one launch
    vxf, vyf, mag, ang = sq(g["vx"]), sq(g["vy"]), sq(g["mag"]), sq(g["ang"])
    valid = sq(vm["valid"]).view(torch.bool)
    if as_np:
        eng.synchronize()
        return tuple(t.cpu().numpy() for t in (vxf, vyf, mag, ang, valid))
    return vxf, vyf, mag, ang, valid


# ---------------------------------------------------------------------------------------------
# clustering (main.py:231-259, 402-434)
# ---------------------------------------------------------------------------------------------
def dbscan_clustering(vx_filtered, vy_filtered, valid_mask, eps=1.0, min_samples=5, engine=None):
    """main.py:231-259 -> (labels intp (n,), valid_indices int64 (n,2)); labels equal
    sklearn's, numbering included.  Raises ValueError on an empty mask like sklearn does
    (the reference's per-pair try/except then skips the pair)."""
    eng = engine or default_engine()
    as_np = _is_np(vx_filtered, vy_filtered, valid_mask)
    vx = _to_dev(eng, vx_filtered)
    vy = _to_dev(eng, vy_filtered)
    # the reference's filtered velocities are float32 values held in float64 (main.py:600-601): narrowed on the
    # device with a check, one launch each
    vx = eng.narrow_f64(vx) if vx.dtype == torch.float64 else vx.to(torch.float32)
    vy = eng.narrow_f64(vy) if vy.dtype == torch.float64 else vy.to(torch.float32)
    valid = _to_dev(eng, valid_mask)
    valid = valid.view(torch.uint8) if valid.dtype == torch.bool else valid.to(torch.uint8)
    n_valid, labels, indices, _ = eng.dbscan_grid(vx, vy, valid, eps, min_samples)
    eng.synchronize()
    n = int(n_valid[0].item())
    if n == 0:
        raise ValueError("Found array with 0 sample(s) (shape=(0, 4)) while a minimum of 1 is required by DBSCAN.")
    lab, idx = labels[0, :n], indices[0, :n]
    if as_np:
        return lab.cpu().numpy().astype(np.intp), idx.cpu().numpy().astype(np.int64)
    return lab.to(torch.int64), idx.to(torch.int64)


def _eigvals_2x2(crr, crc, ccc):
    t = crr + ccc
    d = crr * ccc - crc * crc
    disc = np.sqrt(np.maximum(t * t / 4 - d, 0.0))
    return np.stack([t / 2 + disc, t / 2 - disc], axis=-1)


def extract_cluster_data(labels, indices, vx, vy, engine=None):
    """main.py:402-434 -> {label: {'centroid', 'measurement', 'eigenvalues'}}."""
    eng = engine or default_engine()
    labels_np = labels.cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
    if len(labels_np) != len(indices):
        raise ValueError("Mismatch between labels and valid_indices dimensions.")
    n = len(labels_np)
    if n == 0:
        return {}
    n_clusters = int(labels_np.max()) + 1
    if n_clusters <= 0:
        return {}
    lab = _to_dev(eng, labels, torch.int32).reshape(1, n).contiguous()
    idx = _to_dev(eng, indices, torch.int32).reshape(1, n, 2).contiguous()
    vxd = _to_dev(eng, vx, torch.float32)
    vyd = _to_dev(eng, vy, torch.float32)
    H, W = vxd.shape[-2:]
    idx_np = idx[0].cpu().numpy()
    if np.any(idx_np[:, 0] >= H) or np.any(idx_np[:, 1] >= W):
        raise IndexError("Cluster points are out of bounds for velocity grid.")
    nv = torch.tensor([n], dtype=torch.int32, device=eng.tdev)
    s = eng.cluster_summary(vxd, vyd, nv, lab, idx, n_clusters)
    eng.synchronize()
    s = s[0].cpu().numpy()
    if np.any(s[:, 0] == 1):
        # np.cov of one point is NaN and the reference's np.linalg.eigvals raises on it (main.py:424)
        raise np.linalg.LinAlgError("Array must not contain infs or NaNs")
    eig = _eigvals_2x2(s[:, 5], s[:, 6], s[:, 7])
    out = {}
    for lab_id in range(n_clusters):
        if s[lab_id, 0] <= 0:
            continue
        centroid = np.array([s[lab_id, 1], s[lab_id, 2]])
        out[lab_id] = {"centroid": centroid,
                       "measurement": [centroid[0], centroid[1], s[lab_id, 3], s[lab_id, 4]],
                       "eigenvalues": eig[lab_id]}
    return out


def clusters_from_summary(summary, n_clusters: int, max_clusters: int | None = None) -> dict:
    """{label: {'centroid', 'measurement', 'eigenvalues'}} (extract_cluster_data's result, main.py:402-434) from
    the device's per-cluster rows [count, mean row, mean col, mean vx, mean vy, cov rr, rc, cc].
    Two behaviours of the reference are kept: a cluster of ONE cell has a NaN covariance, on which
    np.linalg.eigvals raises LinAlgError — the reference's per-pair try/except then skips the pair
    (main.py:424, 635-637); and clusters are never dropped silently: more clusters than summary rows raise."""
    if max_clusters is not None and n_clusters > max_clusters:
        raise RuntimeError(f"{n_clusters} clusters but only {max_clusters} summary rows: raise max_clusters")
    s = np.asarray(summary)[:n_clusters]
    if len(s) and np.any(s[:, 0] == 1):
        raise np.linalg.LinAlgError("Array must not contain infs or NaNs")
    eig = _eigvals_2x2(s[:, 5], s[:, 6], s[:, 7]) if len(s) else np.zeros((0, 2))
    return {i: {"centroid": np.array([s[i, 1], s[i, 2]]),
                "measurement": [s[i, 1], s[i, 2], s[i, 3], s[i, 4]],
                "eigenvalues": eig[i]} for i in range(len(s)) if s[i, 0] > 0}


def flow_to_clusters(bev1, bev2, x_range, y_range, dt, alpha_cont, eps, min_samples, farneback=None, engine=None,
                     max_clusters=4096, return_grids=False):
    """The body of the reference's driver loop from the two BEVs to the cluster
    dictionary the EKF consumes (main.py:577-615), as one device-resident chain.
    return_grids adds a 4th result: the filtered velocity grids, their magnitude and curl (f64, as the
    reference holds them, main.py:600-606) for the artefact writers."""
    eng = engine or default_engine()
    a, b = _bev_to_dev(eng, bev1), _bev_to_dev(eng, bev2)
    H, W = a.shape[-2:]
    px = (x_range[1] - x_range[0]) / W
    py = (y_range[1] - y_range[0]) / H
    res = eng.flow_pipeline(a, b, px, py, alpha_cont, eps, min_samples, farneback_params(**(farneback or {})),
                            max_clusters=max_clusters)
    eng.synchronize()
    n = int(res.n_valid[0].item())
    ncl = int(res.n_clusters[0].item())
    labels = res.labels[0, :n].cpu().numpy().astype(np.intp)
    indices = res.indices[0, :n].cpu().numpy().astype(np.int64)
    clusters = clusters_from_summary(res.summary[0, :min(ncl, max_clusters)].cpu().numpy(), ncl, max_clusters)
    if return_grids:
        # f64 like the reference's arrays (f32 * int64 mask), magnitude and f64 curl: one device launch
        g = eng.filtered_grids_f64(res.vx_f[0], res.vy_f[0])
        eng.synchronize()
        grids = dict(vx_filtered=g["vx"].cpu().numpy(), vy_filtered=g["vy"].cpu().numpy(),
                     velocity_magnitude=g["mag"].cpu().numpy(), angular_velocity=g["ang"].cpu().numpy())
        return labels, indices, clusters, grids
    return labels, indices, clusters
