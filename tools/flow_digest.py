#!/usr/bin/env python
"""SHA-256 of the Farneback flow of a few seeded geometries: run it under two settings of an A/B switch
(e.g. DATMO_XM_EDGE_FILL=0 / 1) and compare the lines — bit-identity of a kernel variant."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from datmo_using_optical_flow_b200 import synth  # noqa: E402
from datmo_using_optical_flow_b200.engine import Engine, farneback_params  # noqa: E402

eng = Engine(0)
for H, W, B, kw in [(1024, 1024, 4, {}), (400, 400, 3, {}), (307, 307, 2, {}), (96, 160, 2, {}), (333, 501, 2, {}),
                    (800, 800, 2, dict(pyr_scale=0.5, levels=5)), (64, 2048, 1, {})]:
    a, b = synth.bev_pairs(7, B, H, W)
    out = eng.farneback(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), farneback_params(**kw))
    eng.synchronize()
    print(H, W, B, hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16])
