"""Restatement of the reference's DBSCAN stage — TEST ORACLE.

Follows /root/reference/Optical_flow/main.py:231-259 (``dbscan_clustering``):
features = (row, col, vx_f, vy_f) float64 of the valid cells in row-major
order, ``sklearn.cluster.DBSCAN(eps, min_samples).fit`` (scikit-learn is
unpinned by the reference; 1.9.0 in this image).

``dbscan_clustering_sklearn`` is the reference's own call sequence.
``dbscan_grid`` restates sklearn's result as the grid rule the CUDA kernel
implements, following sklearn/cluster/_dbscan.py:427-462 and
sklearn/cluster/_dbscan_inner.pyx:18-41:
  * neighbours (self included) are the valid cells with
    d2 = drow^2 + dcol^2 + dvx^2 + dvy^2 <= eps^2, fp64, summed in that order;
  * core  <=>  neighbour count >= min_samples;
  * clusters = connected components of core cells under the neighbour relation,
    numbered by ascending minimum row-major rank of their core cells;
  * a non-core cell with a core neighbour takes the lowest-numbered such
    cluster (the DFS of the first cluster reaches it first); otherwise -1.
Pinned against live sklearn in tests/test_oracle_stages.py and against golden
labels produced through the reference's ``dbscan_clustering``.
"""
from __future__ import annotations

import numpy as np


def dbscan_clustering_sklearn(vx_filtered, vy_filtered, valid_mask, eps=1.0, min_samples=5):
    """The reference's call sequence, main.py:246-259."""
    from sklearn.cluster import DBSCAN
    valid_indices = np.array(np.nonzero(valid_mask)).T
    valid_vx = vx_filtered[valid_mask]
    valid_vy = vy_filtered[valid_mask]
    features = np.column_stack((valid_indices, valid_vx, valid_vy))
    if len(features) == 0:
        # sklearn raises on an empty array; the reference's try/except swallows it
        raise ValueError("Found array with 0 sample(s)")
    clustering = DBSCAN(eps=eps, min_samples=min_samples).fit(features)
    return clustering.labels_, valid_indices


def dbscan_grid(vx_filtered, vy_filtered, valid_mask, eps=1.0, min_samples=5):
    """Grid-rule restatement -> (labels intp[n], valid_indices int64[n,2])."""
    valid_mask = np.asarray(valid_mask, dtype=bool)
    H, W = valid_mask.shape
    vx = np.asarray(vx_filtered, dtype=np.float64)
    vy = np.asarray(vy_filtered, dtype=np.float64)
    idx = np.array(np.nonzero(valid_mask)).T.astype(np.int64)
    n = len(idx)
    rank = -np.ones((H, W), dtype=np.int64)
    rank[valid_mask] = np.arange(n)
    r = int(np.floor(eps))
    eps2 = float(eps) * float(eps)
    # neighbour lists over the (2r+1)^2 window
    nbrs = [[] for _ in range(n)]
    rows, cols = idx[:, 0], idx[:, 1]
    for dr in range(-r, r + 1):
        for dc in range(-r, r + 1):
            rr = rows + dr
            cc = cols + dc
            ok = (rr >= 0) & (rr < H) & (cc >= 0) & (cc < W)
            rrc = np.where(ok, rr, 0)
            ccc = np.where(ok, cc, 0)
            ok &= valid_mask[rrc, ccc]
            dvx = vx[rows, cols] - vx[rrc, ccc]
            dvy = vy[rows, cols] - vy[rrc, ccc]
            d2 = float(dr * dr)
            d2 = d2 + float(dc * dc)
            d2 = d2 + dvx * dvx
            d2 = d2 + dvy * dvy
            ok &= d2 <= eps2
            src = np.nonzero(ok)[0]
            dst = rank[rrc[src], ccc[src]]
            for s, d in zip(src.tolist(), dst.tolist()):
                nbrs[s].append(d)
    count = np.array([len(x) for x in nbrs], dtype=np.int64)
    core = count >= min_samples
    # union-find over core cells, root = minimum rank
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for s in range(n):
        if not core[s]:
            continue
        for d in nbrs[s]:
            if core[d]:
                ra, rb = find(s), find(d)
                if ra != rb:
                    if ra < rb:
                        parent[rb] = ra
                    else:
                        parent[ra] = rb
    root = np.array([find(s) if core[s] else -1 for s in range(n)], dtype=np.int64)
    roots = np.unique(root[root >= 0])
    label_of_root = {int(rt): i for i, rt in enumerate(roots.tolist())}
    labels = -np.ones(n, dtype=np.intp)
    for s in range(n):
        if core[s]:
            labels[s] = label_of_root[int(root[s])]
        else:
            best = -1
            for d in nbrs[s]:
                if core[d]:
                    rt = int(root[d])
                    if best < 0 or rt < best:
                        best = rt
            if best >= 0:
                labels[s] = label_of_root[best]
    return labels, idx


def same_partition(a: np.ndarray, b: np.ndarray) -> bool:
    """Partition equivalence up to label permutation; noise (-1) must match exactly."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    if not np.array_equal(a == -1, b == -1):
        return False
    fwd, bwd = {}, {}
    for x, y in zip(a.tolist(), b.tolist()):
        if fwd.setdefault(x, y) != y or bwd.setdefault(y, x) != x:
            return False
    return True
