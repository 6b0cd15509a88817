"""Sequence driver: the reference's ``process_multiple_frames`` loop
(/root/reference/Optical_flow/main.py:541-641) with the hot path on the GPU.

Differences from the reference loop, all deliberate:
  * lines 588-589 of main.py (the unguarded acceleration that raises on the first pair
    and makes every pair fail) are not reproduced; ax / ay were dead values anyway;
  * each frame is preprocessed once and reused as bev2 of pair i-1 and bev1 of pair i
    (the reference preprocesses it twice with different expansion noise);
  * the per-pair results are returned; with ``output_dir`` the data files the reference saves
    (.npy grids, DBSCAN arrays, track YAML, the two CSVs; artefacts.py) are written too — the
    matplotlib PNGs are out of scope.
Per-pair failures are caught and the pair skipped, like the reference's try/except.
"""
from __future__ import annotations

from . import artefacts
from . import main as ops
from .engine import default_engine
from .tracker import TrackManager

# the reference's config.yaml keys that the loop reads (main.py:542-550)
DEFAULT_CONFIG = dict(grid_resolution=[0.2, 0.2], x_range=[-20, 20], y_range=[-20, 20], z_max=2.0,
                      roi_bounds=[-10, 10, -10, 10, -3, 1], masks=dict(alpha_p=[0.8], alpha_cont=[0.2]), dt=1.0,
                      dbscan_params=dict(eps=5.0, min_samples=3))


def process_clouds(clouds, config=None, engine=None, seed=0, ground_masks=None, verbose=False, output_dir=None,
                   save_grids=False):
    """clouds: iterable of float32 (N,4) sweeps of ONE sequence (or .pcd paths).
    Returns dict(tracks=TrackManager, pairs=[per-pair dict], bevs=[uint8 grids or None]).
    output_dir: also write the reference's per-frame files there (saving_utils.py formats);
    save_grids adds the filtered velocity grids and the per-cell CSV (large)."""
    cfg = dict(DEFAULT_CONFIG)
    cfg.update(config or {})
    eng = engine or default_engine()
    alpha_cont = cfg["masks"]["alpha_cont"][0]
    dbp = cfg["dbscan_params"]
    tm = TrackManager()
    bevs, pairs = [], []
    prev = None
    if output_dir is not None:
        import os
        os.makedirs(output_dir, exist_ok=True)
        tracks_csv = os.path.join(output_dir, "tracks.csv")
        cells_csv = os.path.join(output_dir, "filtered_velocities.csv")
        for f in (tracks_csv, cells_csv):
            if os.path.exists(f):
                os.remove(f)   # the reference starts its CSV afresh (main.py:556-558)
    for i, cloud in enumerate(clouds):
        try:
            pts = ops.read_pcd(cloud) if isinstance(cloud, (str, bytes)) else cloud
            gm = None if ground_masks is None else ground_masks[i]
            bev = ops.preprocess_points(pts, cfg["grid_resolution"], cfg["x_range"], cfg["y_range"], cfg["z_max"],
                                        cfg["roi_bounds"], seed=seed + i, ground_mask=gm, engine=eng)
        except Exception as exc:  # the reference prints and moves on (main.py:635-637)
            if verbose:
                print(f"Error preprocessing frame {i}: {exc}")
            bev = None
        bevs.append(bev)
        if output_dir is not None and bev is not None:
            artefacts.save_bev(output_dir, bev, i)
        if i == 0:
            prev = bev
            continue
        rec = dict(index=i - 1, skipped=True)
        if prev is None or bev is None:
            if verbose:
                print(f"Invalid BEV grid for frames {i - 1} and {i}. Skipping.")
        else:
            try:
                out = ops.flow_to_clusters(prev, bev, cfg["x_range"], cfg["y_range"], cfg["dt"], alpha_cont,
                                           dbp["eps"], dbp["min_samples"], engine=eng,
                                           return_grids=output_dir is not None and save_grids)
                labels, indices, clusters = out[:3]
                if len(labels) == 0:
                    raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by DBSCAN.")
                # main.py:618-634: association + EKF, THEN the savers, THEN lifetimes and manage_tracks — a
                # confirmed track that manage_tracks deletes on this pair is still in this pair's files
                tm.associate_and_update(clusters, cfg["dt"])
                saved_tracks = tm.as_array()
                if output_dir is not None:
                    k = i - 1
                    if save_grids:
                        # main.py:600-609: the filtered field, its magnitude and curl (all from the device)
                        g = out[3]
                        artefacts.save_velocity_grid(output_dir, g["vx_filtered"], g["vy_filtered"], k)
                        artefacts.save_all_filtered_velocities_to_csv(g["vx_filtered"], g["vy_filtered"],
                                                                      g["velocity_magnitude"], g["angular_velocity"],
                                                                      k, cells_csv)
                    artefacts.save_dbscan_results(output_dir, labels, indices, k)
                    artefacts.save_ekf_tracks(output_dir, tm, k)
                    artefacts.save_all_velocities_to_csv(tm, k, tracks_csv)
                tm.step_lifetimes()
                rec = dict(index=i - 1, skipped=False, labels=labels, indices=indices, clusters=clusters,
                           tracks=tm.as_array(), saved_tracks=saved_tracks)
            except Exception as exc:
                if verbose:
                    print(f"Error processing frames {i - 1} and {i}: {exc}")
        pairs.append(rec)
        prev = bev
    return dict(tracks=tm, pairs=pairs, bevs=bevs)
