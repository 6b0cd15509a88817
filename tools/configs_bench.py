#!/usr/bin/env python
"""Farneback throughput on the other BASELINE configurations (parity-test cases, not bench lines):
cfg1 400^2, cfg2 800^2 with a 0.5 x 5 pyramid, cfg3 1024^2, cfg4 2048^2 with poly_n 7 / 10 iterations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from datmo_using_optical_flow_b200 import synth  # noqa: E402
from datmo_using_optical_flow_b200.engine import Engine, farneback_params  # noqa: E402

eng = Engine(0)
CFG = [("cfg1 400^2 reference params", 400, 64, {}),
       ("cfg2 800^2 pyr 0.5 x 5", 800, 32, dict(pyr_scale=0.5, levels=5)),
       ("cfg3 1024^2 reference params", 1024, 32, {}),
       ("cfg4 2048^2 poly_n 7, sigma 1.5, 10 iterations", 2048, 8, dict(poly_n=7, poly_sigma=1.5, iterations=10))]
only = os.environ.get("CFG_ONLY")   # e.g. CFG_ONLY=cfg1 (for a launch list of one configuration)
for name, size, B, kw in CFG:
    if only and not name.startswith(only):
        continue
    a, b = synth.bev_pairs(0, B, size, size)
    a, b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    p = farneback_params(**kw)
    out = torch.empty((B, size, size, 2), dtype=torch.float32, device="cuda")
    for _ in range(3):
        eng.farneback(a, b, p, out=out)
    eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    with eng.on_stream():
        e0.record(torch.cuda.current_stream())
        for _ in range(n):
            eng.farneback(a, b, p, out=out)
        e1.record(torch.cuda.current_stream())
    eng.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: {B} pairs in {ms:.2f} ms = {B / ms * 1e3:.0f} pairs/s (layers {eng.farneback_layers(size, size, p)})")
