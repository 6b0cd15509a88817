"""Cell-level emulation of the run-based grid DBSCAN the CUDA library runs — TEST ORACLE.

The product kernel (datmo_using_optical_flow_b200/csrc/dbscan_runs.cu) does not visit
every core-cell pair of the window like oracle/dbscan_np.dbscan_grid does; it works on
ROW RUNS (maximal chains of horizontally adjacent core cells that are pairwise within
eps) and gives every pair of runs within reach of one another to exactly ONE cell, the
"responsible" cell.  This module restates that rule in plain Python so the rule itself
can be pinned against live sklearn (the library the reference calls at
/root/reference/Optical_flow/main.py:257) on the CPU, independently of the CUDA code:

  core        as sklearn: >= min_samples valid cells (self included) with
              d2 = drow^2 + dcol^2 + dvx^2 + dvy^2 <= eps^2, fp64, that summation order
  link(x)     core(x) and core(x-1) and within_eps(x, x-1); a run head is a core cell
              without link; union-find nodes are the run heads (root = minimum index)
  reach       run A (row y) and run B (row y-dr, 0 <= dr <= floor(eps)) can only hold a
              neighbour pair if some |xa - xb| <= rp[dr], rp[dr] = max dc with
              dr^2 + dc^2 <= eps^2
  responsible the first cell of A that sees B in its window: the head of A for every
              run already in the head's window, otherwise the cell at b0 - rp[dr].  Only heads
              do work: A's head looks up at the runs in its window; B's head looks DOWN at the
              one run that covers its window's left edge from further left (that run's first
              cell to see B is b0 - rp[dr])
  pair work   skip when both heads already share a root; else test the cell pairs of
              (A, B) until one is within eps, then union
  passes      1: every head points at the run above its first vertical link (core cell whose
              upper neighbour is core and within eps) inside the head's own 32-cell word — no
              atomics; 2 (after a flatten): all pairs within reach, rows 0..floor(eps) above,
              except that a pair two or more rows apart is dropped when an unbroken column of
              vertical links joins the two runs (the row-1 pairs along it make the union).
"""
from __future__ import annotations

import math

import numpy as np


def _within(dr, dc, vx0, vy0, vx1, vy1, e2):
    d2 = float(dr * dr)
    d2 = d2 + float(dc * dc)
    dvx = vx0 - vx1
    dvy = vy0 - vy1
    d2 = d2 + dvx * dvx
    d2 = d2 + dvy * dvy
    return d2 <= e2


def reach_table(eps: float):
    r = int(math.floor(eps))
    e2 = float(eps) * float(eps)
    rp = []
    for dr in range(r + 1):
        dc = 0
        while float(dr * dr + (dc + 1) * (dc + 1)) <= e2:
            dc += 1
        rp.append(dc if float(dr * dr) <= e2 else -1)
    return r, e2, rp


def dbscan_runs(vx_filtered, vy_filtered, valid_mask, eps=1.0, min_samples=5, stats=None):
    """-> (labels intp[n], valid_indices int64[n,2]); must equal dbscan_np.dbscan_grid / sklearn."""
    valid = np.asarray(valid_mask, dtype=bool)
    H, W = valid.shape
    vx = np.asarray(vx_filtered, dtype=np.float64)
    vy = np.asarray(vy_filtered, dtype=np.float64)
    r, e2, rp = reach_table(eps)
    idx = np.array(np.nonzero(valid)).T.astype(np.int64)
    n = len(idx)
    rank = -np.ones((H, W), dtype=np.int64)
    rank[valid] = np.arange(n)

    def within(y, x, yy, xx):
        return _within(y - yy, x - xx, vx[y, x], vy[y, x], vx[yy, xx], vy[yy, xx], e2)

    # core cells: rows nearest first, early exit (same set as a full count)
    core = np.zeros((H, W), dtype=bool)
    order = [0]
    for d in range(1, r + 1):
        order += [-d, d]
    for y, x in idx.tolist():
        cnt = 0
        for dr in order:
            yy = y + dr
            if yy < 0 or yy >= H:
                continue
            w = rp[abs(dr)]
            for xx in range(max(0, x - w), min(W - 1, x + w) + 1):
                if valid[yy, xx] and within(y, x, yy, xx):
                    cnt += 1
                    if cnt >= min_samples:
                        break
            if cnt >= min_samples:
                break
        core[y, x] = cnt >= min_samples
    link = np.zeros((H, W), dtype=bool)
    for y, x in idx.tolist():
        if x > 0 and core[y, x] and core[y, x - 1] and within(y, x, y, x - 1):
            link[y, x] = True
    head = core & ~link
    parent = {}
    for y, x in np.array(np.nonzero(head)).T.tolist():
        parent[y * W + x] = y * W + x

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    def union(a, b):
        a, b = find(a), find(b)
        if a == b:
            return
        if a < b:
            a, b = b, a
        parent[a] = b

    def head_of(y, x):
        while not head[y, x]:
            x -= 1
        return x

    def run_end(y, x):
        x += 1
        while x < W and core[y, x] and not head[y, x]:
            x += 1
        return x - 1

    counters = dict(pairs=0, skipped=0, tests=0, unions=0)

    def process(y, x, yy, xb_any):
        """run of (y, x) against the run holding (yy, xb_any); (y, x) is the responsible cell."""
        counters["pairs"] += 1
        a0 = head_of(y, x)
        b0 = head_of(yy, xb_any)
        if find(y * W + a0) == find(yy * W + b0):
            counters["skipped"] += 1
            return
        a1, b1 = run_end(y, x), run_end(yy, b0)
        dr = y - yy
        w = rp[dr]
        for xa in range(x, a1 + 1):
            lo = max(b0, xa - w)
            hi = min(b1, xa + w) if dr > 0 else min(b1, xa - 1)
            if xa - w > b1:
                break
            for xb in range(lo, hi + 1):
                counters["tests"] += 1
                if within(y, xa, yy, xb):
                    counters["unions"] += 1
                    union(y * W + a0, yy * W + b0)
                    return

    core_cells = np.array(np.nonzero(core)).T.tolist()
    # vertical links: a core cell whose upper neighbour is core and within eps
    up = np.zeros((H, W), dtype=bool)
    for y, x in core_cells:
        if y > 0 and core[y - 1, x] and within(y, x, y - 1, x):
            up[y, x] = True

    def word_part(y, a0):
        """cells of the run headed at (y, a0) that lie in the head's own 32-cell word"""
        return range(a0, min(run_end(y, a0), (a0 // 32) * 32 + 31) + 1)

    # pass 1 (no atomics in the kernel): every head points at the run above its first vertical link
    for y, a0 in np.array(np.nonzero(head)).T.tolist():
        for x in word_part(y, a0):
            if up[y, x]:
                parent[y * W + a0] = (y - 1) * W + head_of(y - 1, x)
                counters["unions"] += 1
                break
    # pass 2: every pair of runs within reach, rows 0 .. r above.  Rows >= 2 above: the run reached
    # from the head's word part through an unbroken column of vertical links is joined by the row-1
    # pairs along that column, so the pair is dropped without a look at the forest.
    for y, x in core_cells:
        if not head[y, x]:
            continue
        for dr in range(0, r + 1):
            yy = y - dr
            w = rp[dr]
            # looking down: the run of row y + dr that covers the window's left edge with its head further
            # left does not see this head's run from ITS head; its first cell that does is x - w
            yd = y + dr
            if dr > 0 and yd < H and x - w >= 0 and core[yd, x - w] and not head[yd, x - w]:
                process(yd, x - w, y, x)
            if yy < 0:
                continue
            if True:
                implied = -1
                if dr >= 2:
                    for xc in word_part(y, x):
                        if all(up[y - i, xc] for i in range(dr)):
                            implied = head_of(yy, xc)
                            break
                c_lo, c_hi = x - w, (x + w if dr > 0 else x - 1)
                first = True
                for c in range(c_lo, c_hi + 1):
                    if c < 0 or c >= W:
                        first = True   # window clipped by the image edge: the next core cell starts a run
                        continue
                    if core[yy, c] and (first or head[yy, c]):
                        if implied >= 0 and head_of(yy, c) == implied:
                            counters["skipped"] += 1
                        else:
                            process(y, x, yy, c)
                    first = not core[yy, c]
    # labels
    root = -np.ones((H, W), dtype=np.int64)
    for y, x in core_cells:
        root[y, x] = find(y * W + head_of(y, x))
    roots = np.unique(root[root >= 0])
    labels = -np.ones(n, dtype=np.intp)
    for s, (y, x) in enumerate(idx.tolist()):
        if core[y, x]:
            labels[s] = np.searchsorted(roots, root[y, x])
            continue
        best = -1
        for dr in range(-r, r + 1):
            yy = y + dr
            if yy < 0 or yy >= H:
                continue
            w = rp[abs(dr)]
            for xx in range(max(0, x - w), min(W - 1, x + w) + 1):
                if core[yy, xx] and within(y, x, yy, xx):
                    rt = root[yy, xx]
                    if best < 0 or rt < best:
                        best = rt
        if best >= 0:
            labels[s] = np.searchsorted(roots, best)
    if stats is not None:
        stats.update(counters, runs=int(head.sum()), core=int(core.sum()))
    return labels, idx
