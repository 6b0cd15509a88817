"""Generates tests/golden/*.npz by running the UNMODIFIED reference functions
(/root/reference/Optical_flow/main.py, imported through oracle/ref_loader.py with
stubbed open3d / matplotlib / shapely) on small seeded inputs.

Run in the build container only (the reference is not on the GPU box):
    python tests/golden/make_golden.py
Library versions used are recorded inside each file.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from datmo_using_optical_flow_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def versions():
    import cv2
    import sklearn
    return np.array([f"numpy {np.__version__}", f"cv2 {cv2.__version__}", f"sklearn {sklearn.__version__}"])


def bev_cases():
    rng = np.random.default_rng(11)
    cases = {}
    res, xr, yr = (0.5, 0.5), (-10.0, 10.0), (-8.0, 8.0)
    # (a) ordinary cloud, z > 0, some points outside the grid and in the (lo - w, lo) sliver
    p = np.column_stack([rng.uniform(-11, 11, 6000), rng.uniform(-9, 9, 6000), rng.uniform(0.05, 3.0, 6000)])
    cases["a"] = (p, res, xr, yr, 5.0)
    # (b) roof-mounted sensor: most z < 0 (uint8 wrap-around), x10 expansion noise
    base = np.column_stack([rng.uniform(-9, 9, 500), rng.uniform(-7, 7, 500), rng.uniform(-2.4, 0.8, 500)])
    p = np.repeat(base, 10, axis=0) + rng.normal(scale=0.01, size=(5000, 3))
    cases["b"] = (p, res, xr, yr, 2.0)
    # (c) every occupied cell <= 0 -> max = 0 -> all zeros
    p = np.column_stack([rng.uniform(-9, 9, 800), rng.uniform(-7, 7, 800), np.full(800, -1.25)])
    cases["c"] = (p, res, xr, yr, 2.0)
    # (d) non-square resolution, single point, points exactly on bin edges
    p = np.array([[0.0, 0.0, 1.0], [-10.0, -8.0, 0.5], [9.999, 7.999, 2.0], [-10.2, 0.0, 3.0], [10.0, 0.0, 3.0],
                  [0.25, 0.2, 0.7], [0.25, 0.2, 0.9]])
    cases["d"] = (p, (0.25, 0.2), xr, yr, 5.0)
    return cases


def main():
    ref = ref_loader.load_reference_main()
    ver = versions()

    # ---- BEV ---------------------------------------------------------------------------
    out = {"versions": ver}
    for name, (p, res, xr, yr, hmax) in bev_cases().items():
        with ref_loader.quiet(), np.errstate(all="ignore"):
            g = ref.compute_bev_grid(p, list(res), list(xr), list(yr), h_max=hmax)
        out[f"{name}_points"] = p
        out[f"{name}_params"] = np.array([res[0], res[1], xr[0], xr[1], yr[0], yr[1], hmax])
        out[f"{name}_bev"] = g
    pts = np.array([[-11.0, 0, 0], [-10.0, -10.0, -3.0], [10.0, 10.0, 1.0], [0, 0, 1.0000001], [0, 0, 0.5], [3, 11, 0]])
    out["roi_points"] = pts
    out["roi_bounds"] = np.array([-10, 10, -10, 10, -3, 1.0])
    out["roi_out"] = ref.filter_points_in_roi(pts, [-10, 10, -10, 10, -3, 1])
    np.savez_compressed(os.path.join(OUT, "bev.npz"), **out)

    # ---- flow -> velocity -> masks -> dbscan -> clusters, through the reference functions ----
    out = {"versions": ver}
    xr, yr = [-16.0, 16.0], [-12.0, 12.0]
    pairs = {"tex": synth.textured_pair(5, 96, 128, shift=(2, -3)), "blob": synth.bev_pair(3, 120, 160)}
    for name, (a, b) in pairs.items():
        with ref_loader.quiet():
            vx, vy, ang = ref.compute_velocity_vectors(a, b, xr, yr, 1.0)
            mask = ref.continuity_mask(vx, vy, 0.2)
        vx_f = vx * mask
        vy_f = vy * mask
        mag = np.sqrt(vx_f ** 2 + vy_f ** 2)
        dvx_dy, dvx_dx = np.gradient(vx_f)
        dvy_dy, dvy_dx = np.gradient(vy_f)
        ang_f = dvy_dx - dvx_dy
        valid = mag > 0.1
        out[f"{name}_a"], out[f"{name}_b"] = a, b
        out[f"{name}_vx"], out[f"{name}_vy"], out[f"{name}_ang"] = vx, vy, ang
        out[f"{name}_mask"] = mask.astype(np.uint8)
        out[f"{name}_angf"] = ang_f
        out[f"{name}_valid"] = valid
        if valid.sum() >= 12:
            with ref_loader.quiet():
                labels, idx = ref.dbscan_clustering(vx_f, vy_f, valid, eps=5.0, min_samples=3)
                cl = ref.extract_cluster_data(labels, idx, vx_f, vy_f)
            out[f"{name}_labels"], out[f"{name}_indices"] = labels, idx
            keys = sorted(cl)
            out[f"{name}_cl_keys"] = np.array(keys)
            out[f"{name}_cl_meas"] = np.array([cl[k]["measurement"] for k in keys], dtype=np.float64)
            out[f"{name}_cl_eig"] = np.array([np.sort(np.real(cl[k]["eigenvalues"])) for k in keys])
    out["ranges"] = np.array(xr + yr)
    np.savez_compressed(os.path.join(OUT, "flow_chain.npz"), **out)

    # ---- DBSCAN on a crafted field with border points and several (eps, min_samples) ----------
    out = {"versions": ver}
    rng = np.random.default_rng(21)
    H, W = 48, 64
    vx = np.zeros((H, W), np.float32)
    vy = np.zeros((H, W), np.float32)
    for _ in range(9):
        y, x = rng.integers(2, H - 12), rng.integers(2, W - 14)
        h, w = rng.integers(2, 10), rng.integers(2, 12)
        vx[y:y + h, x:x + w] = rng.uniform(-3, 3)
        vy[y:y + h, x:x + w] = rng.uniform(-3, 3)
    vx += (rng.uniform(-0.4, 0.4, (H, W)) * (vx != 0)).astype(np.float32)
    sp = rng.uniform(size=(H, W)) < 0.03
    vx[sp] = rng.uniform(-2, 2, sp.sum()).astype(np.float32)
    vy[sp] = rng.uniform(-2, 2, sp.sum()).astype(np.float32)
    vxd, vyd = vx.astype(np.float64), vy.astype(np.float64)
    valid = np.sqrt(vxd ** 2 + vyd ** 2) > 0.1
    out["vx"], out["vy"], out["valid"] = vx, vy, valid
    for i, (eps, ms) in enumerate([(5.0, 3), (1.0, 5), (1.5, 4), (2.9, 8), (3.0, 2)]):
        with ref_loader.quiet():
            labels, idx = ref.dbscan_clustering(vxd, vyd, valid, eps=eps, min_samples=ms)
        out[f"labels_{i}"] = labels
        out[f"params_{i}"] = np.array([eps, ms])
    out["indices"] = idx
    np.savez_compressed(os.path.join(OUT, "dbscan.npz"), **out)
    for f in ("bev.npz", "flow_chain.npz", "dbscan.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")




def propagation_golden():
    """propagation_mask / propagation_mask_with_acceleration (main.py:166-221) — the reference's own
    double loops — on small seeded fields: f32 velocities as compute_velocity_vectors returns them,
    an f64 case, collisions (several sources per target), sources that leave the grid."""
    ref = ref_loader.load_reference_main()
    rng = np.random.default_rng(31)
    out = {"versions": versions()}
    cases = []
    for k, (h, w, scale, dt, gr, alpha, dtype) in enumerate([
            (40, 56, 0.6, 1.0, (0.25, 0.25), 0.2, np.float32),
            (33, 47, 3.0, 0.5, (0.25, 0.2), 0.5, np.float32),
            (25, 31, 9.0, 0.1, (0.125, 0.125), 1.0, np.float32),
            (21, 18, 2.0, 1.0, (0.25, 0.25), 0.3, np.float64)]):
        # piecewise-constant blobs (targets collide) + noise
        vx = np.zeros((h, w))
        vy = np.zeros((h, w))
        for _ in range(6):
            y, x = rng.integers(0, h - 6), rng.integers(0, w - 6)
            vx[y:y + 6, x:x + 6] = rng.normal() * scale
            vy[y:y + 6, x:x + 6] = rng.normal() * scale
        vx = (vx + rng.normal(scale=0.05 * scale, size=(h, w))).astype(dtype)
        vy = (vy + rng.normal(scale=0.05 * scale, size=(h, w))).astype(dtype)
        ax = rng.normal(scale=scale, size=(h, w)).astype(dtype)
        ay = rng.normal(scale=scale, size=(h, w)).astype(dtype)
        out[f"vx_{k}"], out[f"vy_{k}"], out[f"ax_{k}"], out[f"ay_{k}"] = vx, vy, ax, ay
        out[f"params_{k}"] = np.array([dt, gr[0], gr[1], alpha])
        out[f"mask_{k}"] = ref.propagation_mask(vx, vy, dt, list(gr), alpha)
        out[f"mask_acc_{k}"] = ref.propagation_mask_with_acceleration(vx, vy, ax, ay, dt, list(gr), alpha)
        cases.append(k)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "propagation.npz"), **out)
    print("propagation.npz", os.path.getsize(os.path.join(OUT, "propagation.npz")), "bytes",
          [float(out[f"mask_{k}"].mean()) for k in cases])


def tracker_golden():
    """Reference EKF / track_clusters / manage_tracks + the driver's lifetime bookkeeping
    (main.py:437-515, 618-634) on a seeded stream of cluster dictionaries."""
    ref = ref_loader.load_reference_main()
    rng = np.random.default_rng(77)
    frames = []
    centers = rng.uniform(20, 180, (4, 2))
    vels = rng.uniform(-0.3, 0.3, (4, 2))
    for f in range(24):
        cl = {}
        k = 0
        for j in range(4):
            if rng.uniform() < 0.15:
                continue                                   # missed detection
            c = centers[j] + vels[j] * f * (0.2 if j < 2 else 3.0) + rng.normal(0, 0.05, 2)
            eig = np.abs(rng.normal(0.05, 0.02, 2)) if j < 3 else np.abs(rng.normal(4.0, 1.0, 2))
            cl[k] = dict(centroid=c, measurement=[c[0], c[1], vels[j][0], vels[j][1]], eigenvalues=eig)
            k += 1
        # a parked object seen on every pair: its track is matched every time, gets confirmed, and
        # manage_tracks deletes it once its lifetime passes N2 = 15 (main.py:505-507)
        c = np.array([100.0, 60.0]) + rng.normal(0, 0.01, 2)
        cl[k] = dict(centroid=c, measurement=[c[0], c[1], 0.0, 0.0], eigenvalues=np.abs(rng.normal(0.02, 0.005, 2)))
        frames.append(cl)
    tracks, lifetimes, confirmed = {}, {}, set()
    out = {"versions": versions(), "n_frames": np.array(len(frames))}
    for f, cl in enumerate(frames):
        keys = sorted(cl)
        out[f"f{f}_centroid"] = np.array([cl[k]["centroid"] for k in keys]).reshape(-1, 2)
        out[f"f{f}_meas"] = np.array([cl[k]["measurement"] for k in keys]).reshape(-1, 4)
        out[f"f{f}_eig"] = np.array([cl[k]["eigenvalues"] for k in keys]).reshape(-1, 2)
        tracks = ref.track_clusters(tracks, cl, 1.0, np.eye(4) * 0.1, np.eye(4) * 0.05, gamma=0.5)
        # what save_ekf_tracks / save_all_velocities_to_csv see (main.py:619-620): the table BEFORE the
        # lifetime update and manage_tracks
        out[f"f{f}_saved"] = np.array([[tid, *tracks[tid].state.tolist()] for tid in tracks],
                                      dtype=np.float64).reshape(-1, 5)
        for tid in list(lifetimes.keys()):
            if tid in tracks:
                lifetimes[tid] += 1
            else:
                del lifetimes[tid]
        for tid in tracks:
            if tid not in lifetimes:
                lifetimes[tid] = 1
        ref.manage_tracks(tracks, lifetimes, confirmed, M1=1, N1=4, M2=10, N2=15)
        rows = [[tid, *tracks[tid].state.tolist(), float(tid in confirmed)] for tid in sorted(tracks)]
        out[f"f{f}_tracks"] = np.array(rows, dtype=np.float64).reshape(-1, 6)
    np.savez_compressed(os.path.join(OUT, "tracks.npz"), **out)
    print("tracks.npz", os.path.getsize(os.path.join(OUT, "tracks.npz")), "bytes")


if __name__ == "__main__":
    if "--propagation-only" in sys.argv:
        propagation_golden()
    else:
        if "--tracks-only" not in sys.argv:
            main()
            propagation_golden()
        tracker_golden()
