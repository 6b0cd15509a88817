"""Restatement of ``extract_cluster_data`` — TEST ORACLE.

Follows /root/reference/Optical_flow/main.py:402-434: per label (noise
skipped) centroid = mean (row, col), velocity = mean (vx, vy) at the member
cells, ``np.cov`` (ddof 1) of the member coordinates and its eigenvalues.
``np.linalg.eigvals`` of the symmetric 2x2 returns the eigenvalues in LAPACK's
order; the closed form below returns them as (lambda_a, lambda_b) = the pair
{(t +- sqrt(t^2-4d))/2}; comparisons in the tests are order-insensitive.
"""
from __future__ import annotations

import numpy as np


def extract_cluster_data(labels, indices, vx, vy):
    """Returns {label: dict(centroid, measurement, eigenvalues, count)}."""
    labels = np.asarray(labels)
    indices = np.asarray(indices)
    if len(labels) != len(indices):
        raise ValueError("Mismatch between labels and valid_indices dimensions.")
    out = {}
    for lab in np.unique(labels):
        if lab == -1:
            continue
        pts = indices[labels == lab]
        cvx = vx[pts[:, 0], pts[:, 1]]
        cvy = vy[pts[:, 0], pts[:, 1]]
        centroid = np.mean(pts, axis=0)
        vel = [np.mean(cvx), np.mean(cvy)]
        with np.errstate(all="ignore"):
            cov = np.atleast_2d(np.cov(pts.T))
            if cov.shape == (2, 2) and np.all(np.isfinite(cov)):
                t = cov[0, 0] + cov[1, 1]
                d = cov[0, 0] * cov[1, 1] - cov[0, 1] * cov[1, 0]
                disc = np.sqrt(max(t * t / 4 - d, 0.0))
                eig = np.array([t / 2 + disc, t / 2 - disc])
            else:
                eig = np.array([np.nan, np.nan])
        out[int(lab)] = dict(centroid=centroid,
                             measurement=[centroid[0], centroid[1], vel[0], vel[1]],
                             eigenvalues=eig, count=len(pts))
    return out
