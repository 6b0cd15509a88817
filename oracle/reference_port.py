"""The reference's stage functions as a travelling CPU port — TEST ORACLE / CPU BASELINE.

/root/reference does not exist on the GPU box, so this restates the call
sequence of /root/reference/Optical_flow/main.py using the same third-party
calls the reference makes (cv2.calcOpticalFlowFarneback, sklearn DBSCAN,
numpy).  It is what ``bench.py`` times as ``cpu_baseline`` (kind "port") and
what ``bench.py --impl reference`` runs.  Differences from main.py: the debug
``print`` calls (main.py:141-161) and the file savers are dropped; nothing else.
"""
from __future__ import annotations

import numpy as np

from . import bev_np, cluster_np, dbscan_np, masks_np

# main.py:132-140 — hard-coded, config.yaml's farneback_params block is ignored
FARNEBACK_PARAMS = dict(pyr_scale=0.3, levels=5, winsize=15, iterations=5,
                        poly_n=5, poly_sigma=5, flags=0)


def compute_velocity_vectors(bev1, bev2, x_range, y_range, dt, farneback_params=None):
    """main.py:131-164."""
    import cv2
    p = dict(FARNEBACK_PARAMS)
    if farneback_params:
        p.update(farneback_params)
    flow = cv2.calcOpticalFlowFarneback(bev1.astype(np.float32), bev2.astype(np.float32), None, **p)
    vx, vy = flow[..., 0], flow[..., 1]
    pixel_size_x = (x_range[1] - x_range[0]) / bev1.shape[1]
    pixel_size_y = (y_range[1] - y_range[0]) / bev1.shape[0]
    velocity_x = vx * pixel_size_x
    velocity_y = vy * pixel_size_y
    dvx_dy, dvx_dx = np.gradient(velocity_x)
    dvy_dy, dvy_dx = np.gradient(velocity_y)
    return velocity_x, velocity_y, dvy_dx - dvx_dy


def continuity_mask(vx, vy, alpha_cont):
    """main.py:224-228."""
    div_v = np.gradient(vx, axis=1) + np.gradient(vy, axis=0)
    curl_v = np.gradient(vy, axis=1) - np.gradient(vx, axis=0)
    return ((np.abs(div_v) <= alpha_cont) & (np.abs(curl_v) <= alpha_cont)).astype(int)


def flow_to_clusters(bev1, bev2, x_range, y_range, dt, alpha_cont, eps, min_samples,
                     farneback_params=None, with_clusters=True):
    """One pass of the driver loop body, main.py:577-615, savers removed."""
    vx, vy, _ = compute_velocity_vectors(bev1, bev2, x_range, y_range, dt, farneback_params)
    mask = continuity_mask(vx, vy, alpha_cont)
    vx_f = vx * mask
    vy_f = vy * mask
    mag = np.sqrt(vx_f ** 2 + vy_f ** 2)
    valid = mag > 0.1
    if not valid.any():
        return dict(vx=vx, vy=vy, labels=np.zeros(0, dtype=np.intp),
                    indices=np.zeros((0, 2), dtype=np.int64), clusters={})
    labels, indices = dbscan_np.dbscan_clustering_sklearn(vx_f, vy_f, valid, eps, min_samples)
    clusters = cluster_np.extract_cluster_data(labels, indices, vx_f, vy_f) if with_clusters else {}
    return dict(vx=vx, vy=vy, labels=labels, indices=indices, clusters=clusters)


def preprocess_points(points, grid_resolution, x_range, y_range, z_max, roi_bounds, noise,
                      ground_mask=None, expansion_factor=10):
    """main.py:59-95 minus the PCD read and Open3D RANSAC (the caller passes the
    ground mask): flip x, drop ground, ROI crop, expand with the given noise, rasterise."""
    pts = np.array(points, dtype=np.float64)[:, :3]
    pts[:, 0] = -pts[:, 0]
    if ground_mask is not None:
        pts = pts[~ground_mask]
    roi = bev_np.filter_points_in_roi(pts, roi_bounds)
    if roi.size == 0:
        return None
    exp = bev_np.increase_point_density(roi, expansion_factor, 0.01, noise=noise)
    return bev_np.compute_bev_grid(exp, grid_resolution, x_range, y_range, h_max=z_max)
