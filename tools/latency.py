#!/usr/bin/env python
"""Per-pair latency of the flow -> clusters chain for ONE pair (the sequence driver's regime):
wall clock per call against summed device time, to see how launch-bound small grids are."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from datmo_using_optical_flow_b200 import synth  # noqa: E402
from datmo_using_optical_flow_b200.engine import Engine, farneback_params  # noqa: E402

eng = Engine(0)
for size, B in [(400, 1), (800, 1), (1024, 1), (400, 8)]:
    a, b = synth.bev_pairs(0, B, size, size)
    a, b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    p = farneback_params()
    run = lambda: eng.flow_pipeline(a, b, 0.25, 0.25, 0.2, 5.0, 3, p, cap=size * size // 2, max_clusters=1024, keep_flow=False)
    for _ in range(5):
        run()
    eng.synchronize()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        run()
    eng.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    eng.profile(True)
    eng.profile_reset()
    l0 = eng.launch_count()
    for _ in range(n):
        run()
    pr = eng.profile_read()
    eng.profile(False)
    dev = sum(v["ms"] for v in pr.values()) / n
    print(f"{size}x{size} B={B}: wall {wall:.3f} ms/call, kernels {dev:.3f} ms/call, {(eng.launch_count() - l0) // n} launches/call")

# the same through the C-ABI chain (host frames in, host results out), with its launch train replayed as a
# CUDA graph and launched plainly
from datmo_using_optical_flow_b200.engine import HostFlowPipeline  # noqa: E402

for graph in ("1", "0"):
    os.environ["DATMO_CHAIN_GRAPH"] = graph
    for size, B in [(400, 1), (1024, 1), (400, 8), (1024, 8)]:
        a, b = synth.bev_pairs(0, B, size, size)
        a, b = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
        pipe = HostFlowPipeline(eng, B, size, size, 0.25, 0.25, 0.2, 5.0, 3, farneback_params(), cap=size * size // 2,
                                max_clusters=1024, n_slots=1, want_cells=False)
        for _ in range(5):
            pipe.submit(0, a, b)
            pipe.collect(0)
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            pipe.submit(0, a, b)
            pipe.collect(0)
        wall = (time.perf_counter() - t0) / n * 1e3
        print(f"chain {size}x{size} B={B} graph={graph}: {wall:.3f} ms per submit + collect (summaries only)")
        pipe.close()
