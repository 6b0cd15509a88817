"""Sequence driver: the reference's ``process_multiple_frames`` loop
(/root/reference/Optical_flow/main.py:541-641) with the hot path on the GPU.

Differences from the reference loop, all deliberate:
  * lines 588-589 of main.py (the unguarded acceleration that raises on the first pair
    and makes every pair fail) are not reproduced; ax / ay were dead values anyway;
  * each frame is preprocessed once and reused as bev2 of pair i-1 and bev1 of pair i
    (the reference preprocesses it twice with different expansion noise);
  * the matplotlib / CSV / YAML savers are not called (out of scope); the per-pair
    results are returned instead.
Per-pair failures are caught and the pair skipped, like the reference's try/except.
"""
from __future__ import annotations

import numpy as np

from . import main as ops
from .engine import default_engine
from .tracker import TrackManager

# the reference's config.yaml keys that the loop reads (main.py:542-550)
DEFAULT_CONFIG = dict(grid_resolution=[0.2, 0.2], x_range=[-20, 20], y_range=[-20, 20], z_max=2.0,
                      roi_bounds=[-10, 10, -10, 10, -3, 1], masks=dict(alpha_p=[0.8], alpha_cont=[0.2]), dt=1.0,
                      dbscan_params=dict(eps=5.0, min_samples=3))


def process_clouds(clouds, config=None, engine=None, seed=0, ground_masks=None, verbose=False):
    """clouds: iterable of float32 (N,4) sweeps of ONE sequence (or .pcd paths).
    Returns dict(tracks=TrackManager, pairs=[per-pair dict], bevs=[uint8 grids or None])."""
    cfg = dict(DEFAULT_CONFIG)
    cfg.update(config or {})
    eng = engine or default_engine()
    alpha_cont = cfg["masks"]["alpha_cont"][0]
    dbp = cfg["dbscan_params"]
    tm = TrackManager()
    bevs, pairs = [], []
    prev = None
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
        if i == 0:
            prev = bev
            continue
        rec = dict(index=i - 1, skipped=True)
        if prev is None or bev is None:
            if verbose:
                print(f"Invalid BEV grid for frames {i - 1} and {i}. Skipping.")
        else:
            try:
                labels, indices, clusters = ops.flow_to_clusters(prev, bev, cfg["x_range"], cfg["y_range"], cfg["dt"],
                                                                 alpha_cont, dbp["eps"], dbp["min_samples"], engine=eng)
                if len(labels) == 0:
                    raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by DBSCAN.")
                tm.update(clusters, cfg["dt"])
                rec = dict(index=i - 1, skipped=False, labels=labels, indices=indices, clusters=clusters,
                           tracks=tm.as_array())
            except Exception as exc:
                if verbose:
                    print(f"Error processing frames {i - 1} and {i}: {exc}")
        pairs.append(rec)
        prev = bev
    return dict(tracks=tm, pairs=pairs, bevs=bevs)
