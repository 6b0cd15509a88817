"""Where is the reference's own flow well defined? — TEST ORACLE.

``cv2.calcOpticalFlowFarneback`` (the call at /root/reference/Optical_flow/main.py:142) is
ill-conditioned wherever the 15x15 window sees no texture at some pyramid layer: the 2x2 system
is then decided by its 1e-3 regulariser, and an error made at a coarse layer moves the sampling
positions of every finer layer.  On sparse BEV frames cv2's result moves by up to ~1 px when its
INPUT is perturbed by 3e-5 (one float32 ulp of a 0..255 image).  A second implementation cannot
agree with cv2 there better than cv2 agrees with itself, so the parity tolerance of north_star
(max |d flow| <= 1e-3 px, mean <= 1e-5) is asserted on the pixels where cv2 is self-stable, and
the fraction of such pixels is reported.

``self_deviation`` measures it from the reference itself: the per-pixel maximum, over a few
seeded trials, of |flow(perturbed frames) - flow(frames)|.
"""
from __future__ import annotations

import numpy as np

STABLE_TOL = 1e-4      # a pixel is "stable" when no trial moved cv2's own flow by more than this (px)
PERTURBATION = 3e-5    # uniform +- amplitude added to the float32 frames: ~1 ulp at 255
TRIALS = 4


def self_deviation(prev, nxt, cv_params: dict, trials: int = TRIALS, amp: float = PERTURBATION, seed: int = 123):
    """-> (reference flow [H,W,2] f32, per-pixel max self-deviation [H,W] f32)."""
    import cv2
    a = np.asarray(prev, dtype=np.float32)
    b = np.asarray(nxt, dtype=np.float32)
    ref = cv2.calcOpticalFlowFarneback(a, b, None, **cv_params)
    rng = np.random.default_rng(seed)
    s = np.zeros(a.shape, np.float32)
    for _ in range(trials):
        pa = a + rng.uniform(-amp, amp, a.shape).astype(np.float32)
        pb = b + rng.uniform(-amp, amp, b.shape).astype(np.float32)
        s = np.maximum(s, np.abs(cv2.calcOpticalFlowFarneback(pa, pb, None, **cv_params) - ref).max(axis=2))
    return ref, s


def compare(flow, prev, nxt, cv_params: dict, stable_tol: float = STABLE_TOL) -> dict:
    """Deviation of `flow` from cv2 on the frame pair, split by cv2's own stability.  A pixel counts as
    stable when cv2 moved by <= stable_tol in every trial at every pixel of its own winsize x winsize
    window (the flow of a pixel is solved from sums over that window, so an unstable neighbour reaches it)."""
    from scipy.ndimage import maximum_filter
    ref, s = self_deviation(prev, nxt, cv_params)
    d = np.abs(np.asarray(flow) - ref).max(axis=2)
    win = 2 * (int(cv_params.get("winsize", 15)) // 2) + 1
    stable = maximum_filter(s, size=win, mode="nearest") <= stable_tol
    out = dict(stable_fraction=float(stable.mean()), max_stable=float(d[stable].max()) if stable.any() else 0.0,
               mean_stable=float(d[stable].mean()) if stable.any() else 0.0, max_all=float(d.max()),
               mean_all=float(d.mean()), p999_all=float(np.quantile(d, 0.999)),
               ref_self_max=float(s.max()), ref_self_mean=float(s.mean()))
    rest = ~stable
    out["max_unstable"] = float(d[rest].max()) if rest.any() else 0.0
    # on the unstable pixels: how our deviation compares with cv2's own
    out["unstable_ratio"] = float((d[rest] / np.maximum(s[rest], stable_tol)).max()) if rest.any() else 0.0
    return out
