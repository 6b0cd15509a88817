"""Host-side tracking stage: EKF per track, nearest-track gating, M-of-N track management.

Restates the behaviour of the reference's ``EKF`` / ``track_clusters`` /
``manage_tracks`` and the lifetime bookkeeping of its driver loop
(/root/reference/Optical_flow/main.py:437-515, 618-634).  It is O(#clusters)
4x4 scalar work per frame pair and stays on the host; the GPU
hands it the per-cluster summaries.  The reference's quirks are reproduced, not
fixed, because the tracks are part of its observable output:
  * the "state" is [row, col, mean vx, mean vy] but predict() treats state[2] as a
    heading and state[3] as a speed, with the cluster's (vx, vy) as (v, omega);
  * association compares [centroid, eigenvalues] with [x, y, 0, 0] under gamma;
  * tracks not matched in a frame are dropped; every unmatched cluster of one frame
    gets the same new id (max existing id + 1), so only the last one survives;
  * two clusters may update the same track in one frame.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Track:
    state: np.ndarray                      # (4,)
    P: np.ndarray = field(default_factory=lambda: np.eye(4))
    F: np.ndarray = field(default_factory=lambda: np.eye(4))


class TrackManager:
    """tracks / lifetimes / confirmed set of one sequence."""

    def __init__(self, process_noise=None, measurement_noise=None, gamma=0.5, M1=1, N1=4, M2=10, N2=15):
        # main.py:618 and :634 hard-code these
        self.Q = np.eye(4) * 0.1 if process_noise is None else np.asarray(process_noise, dtype=float)
        self.R = np.eye(4) * 0.05 if measurement_noise is None else np.asarray(measurement_noise, dtype=float)
        self.gamma = gamma
        self.M1, self.N1, self.M2, self.N2 = M1, N1, M2, N2
        self.tracks: dict[int, Track] = {}
        self.lifetimes: dict[int, int] = {}
        self.confirmed: set[int] = set()

    # -- EKF ---------------------------------------------------------------------------------
    def _predict(self, t: Track, dt: float, v: float, omega: float):
        theta = t.state[2]
        t.F[0, 2] = dt
        t.F[1, 3] = dt
        speed = t.state[3]
        t.state[0] += speed * math.cos(theta) * dt
        t.state[1] += speed * math.sin(theta) * dt
        t.state[2] += omega * dt
        t.state[3] += v * dt
        t.P = t.F @ t.P @ t.F.T + self.Q

    def _update(self, t: Track, z: np.ndarray):
        # H = I
        innov = z - t.state
        S = t.P + self.R
        K = t.P @ np.linalg.inv(S)
        t.state = t.state + K @ innov
        t.P = (np.eye(4) - K) @ t.P

    # -- association -----------------------------------------------------------------------------
    def associate_and_update(self, clusters: dict, dt: float):
        """clusters: {label: {'centroid', 'measurement', 'eigenvalues'}} in label order."""
        old = self.tracks
        new: dict[int, Track] = {}
        next_id = max(old.keys(), default=0) + 1
        # every track's (row, col) up front: a cluster's distances to all tracks are one vector expression
        # (many tracks) or a loop over plain floats (a handful) instead of one numpy norm per (cluster, track).  A
        # track updated by an earlier cluster of this frame is seen with its new state, as in the reference's
        # loop: its row is refreshed after the update
        ids = list(old.keys())
        few = len(ids) <= 8                      # a handful of tracks: plain float arithmetic beats numpy's call overhead
        pos = [[float(t.state[0]), float(t.state[1])] for t in old.values()]
        if not few:
            pos = np.array(pos, dtype=float).reshape(-1, 2)
        last_new = None      # every unmatched cluster takes the SAME new id: only the last one's track survives
        for _, cl in clusters.items():
            c, ev = cl["centroid"], np.real(cl["eigenvalues"])
            best = None
            # || [centroid, eigenvalues] - [row, col, 0, 0] ||, summed in the order of the 4-vector norm; the first
            # of equal minima wins (the strict < of the reference's loop), a NaN distance never matches
            if few:
                c0, c1, e0, e1 = float(c[0]), float(c[1]), float(ev[0]), float(ev[1])
                best_d = math.inf
                for k, (pr, pc) in enumerate(pos):
                    dr, dc = c0 - pr, c1 - pc
                    d = math.sqrt(dr * dr + dc * dc + e0 * e0 + e1 * e1)
                    if d < best_d and d < self.gamma:
                        best, best_d, i = ids[k], d, k
            elif ids:
                dr, dc = c[0] - pos[:, 0], c[1] - pos[:, 1]
                d = np.sqrt(dr * dr + dc * dc + ev[0] * ev[0] + ev[1] * ev[1])
                d[np.isnan(d)] = np.inf
                i = int(np.argmin(d))
                if d[i] < self.gamma:
                    best = ids[i]
            if best is not None:
                z = np.asarray(cl["measurement"], dtype=float)
                t = old[best]
                self._predict(t, dt, z[2], z[3])
                self._update(t, z)
                new[best] = t
                pos[i][0], pos[i][1] = float(t.state[0]), float(t.state[1])
            else:
                new.setdefault(next_id, None)    # keeps the position of the first unmatched cluster in the dict
                last_new = cl["measurement"]
        if last_new is not None:
            new[next_id] = Track(state=np.array(last_new, dtype=float))
        self.tracks = new
        return new

    # -- lifetimes + M-of-N --------------------------------------------------------------------------
    def step_lifetimes(self):
        for tid in list(self.lifetimes):
            if tid in self.tracks:
                self.lifetimes[tid] += 1
            else:
                del self.lifetimes[tid]
        for tid in self.tracks:
            self.lifetimes.setdefault(tid, 1)
        for tid in list(self.tracks):
            life = self.lifetimes[tid]
            if tid in self.confirmed:
                if life > self.N2 and life - self.M2 <= self.N2:
                    del self.tracks[tid]
            elif life >= self.N1 and life - self.M1 <= self.N1:
                self.confirmed.add(tid)

    def update(self, clusters: dict, dt: float):
        """One frame pair: what main.py:618-634 does between extract_cluster_data and the next pair."""
        self.associate_and_update(clusters, dt)
        self.step_lifetimes()
        return self.tracks

    def as_array(self, max_tracks: int | None = None) -> np.ndarray:
        """(n, 6) float64 rows [track id, row, col, s2, s3, confirmed] sorted by id (padded with NaN)."""
        rows = [[tid, *t.state.tolist(), float(tid in self.confirmed)] for tid, t in sorted(self.tracks.items())]
        arr = np.array(rows, dtype=np.float64).reshape(-1, 6)
        if max_tracks is not None:
            out = np.full((max_tracks, 6), np.nan)
            out[:min(len(arr), max_tracks)] = arr[:max_tracks]
            return out
        return arr
