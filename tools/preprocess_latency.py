#!/usr/bin/env python
"""Device time of the cloud -> BEV stage (flip, RANSAC ground drop, ROI, x10 expansion, rasterise)
for the BASELINE cloud sizes: 32 / 64 / 128 beams, ~60k / 120k / 240k returns."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from datmo_using_optical_flow_b200 import synth  # noqa: E402
from datmo_using_optical_flow_b200.engine import Engine  # noqa: E402

eng = Engine(0)
for beams, n_pts, res, half in [(32, 60_000, 0.25, 50.0), (64, 120_000, 0.125, 50.0), (128, 240_000, 0.1, 51.2)]:
    cloud = torch.from_numpy(synth.lidar_sweep(0, 0, beams, n_pts, 3)).cuda()
    args = ([res, res], [-half, half], [-half, half], 2.0, [-half, half, -half, half, -3, 1])
    for _ in range(3):
        eng.preprocess(cloud, *args, seed=1)
    eng.synchronize()
    n = 20
    t0 = time.perf_counter()
    for i in range(n):
        eng.preprocess(cloud, *args, seed=i)
    eng.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    eng.profile(True)
    eng.profile_reset()
    for i in range(n):
        eng.preprocess(cloud, *args, seed=i)
    pr = eng.profile_read()
    eng.profile(False)
    parts = {k: round(v["ms"] / n, 3) for k, v in pr.items() if v["launches"]}
    print(f"{beams} beams, {cloud.shape[0]} pts, grid {int(2 * half / res)}^2: wall {wall:.3f} ms/frame, device {parts}")
