#!/usr/bin/env python
"""Per-kernel device time of the DBSCAN stage on bench-like inputs (run with DATMO_DBSCAN_SUBTAGS=1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from datmo_using_optical_flow_b200 import synth  # noqa: E402
from datmo_using_optical_flow_b200.engine import Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = Engine(0)
a, b = synth.bev_pairs(0, B, 1024, 1024)
a, b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
flow = eng.farneback(a, b)
vm = eng.velocity_mask(flow, 0.1, 0.1, 0.2, 0.1, want=("vx_f", "vy_f", "valid"))
for _ in range(3):
    eng.dbscan_grid(vm["vx_f"], vm["vy_f"], vm["valid"], 5.0, 3, cap=524288)
eng.profile(True)
eng.profile_reset()
n = 5
for _ in range(n):
    out = eng.dbscan_grid(vm["vx_f"], vm["vy_f"], vm["valid"], 5.0, 3, cap=524288)
p = eng.profile_read()
if os.environ.get("DATMO_DBSCAN_CELLS"):
    names = {"pyramid": "flag scans (2x3 kernels)", "polyexp": "k_core", "flow_init": "k_link_near",
             "flow_iter": "k_flatten (x3)", "velmask": "k_union_far", "dbscan": "k_union_near", "bev": "k_labels"}
else:
    names = {"pyramid": "k_run_pack + 2 x k_run_scan", "polyexp": "k_run_core", "flow_init": "k_run_heads",
             "flow_iter": "k_run_flatten_list (x3)", "velmask": "k_run_pairs (row 0, rows 1..r)",
             "dbscan": "(unused)", "bev": "k_run_labels"}
tot = 0
for k, v in p.items():
    if v["launches"]:
        print(f"{names.get(k, k):28s} {v['ms'] / n:8.3f} ms/call  ({v['launches'] // n} launches)")
        tot += v["ms"] / n
print(f"{'total':28s} {tot:8.3f} ms/call for {B} pairs; valid/pair {out[0].float().mean().item():.0f}, clusters/pair {out[3].float().mean().item():.0f}")
