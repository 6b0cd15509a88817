#!/usr/bin/env python
"""Compact per-kernel summary of an `ncu --set full` report.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep [> profiles/rNN_X.txt]

Reads the report through `ncu -i X --page raw --csv` and prints, per captured launch, the
figures DESIGN.md / bench.py quote: duration, DRAM bytes, pipe and memory throughputs, issue
utilisation, occupancy limiters and the top warp-stall reasons.
"""
import csv
import io
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "L1 data-stage wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe FP64 %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_allocated", "smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
]


def traffic_json(rep: str, note: str) -> None:
    """`--traffic NOTE`: DRAM read + write bytes of the first captured launch as JSON (bench.py's roofline.traffic)."""
    import json
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, first = rows[0], rows[1], rows[2]
    d, u = dict(zip(head, first)), dict(zip(head, units))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def nbytes(k):
        return float(d[k].replace(",", "")) * scale[u[k]]
    t_us = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}[u["gpu__time_duration.sum"]]
    print(json.dumps({"kernel": d["Kernel Name"].split("(")[0], "grid": d["Grid Size"], "block": d["Block Size"],
                      "dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"),
                      "dram_read_bytes": nbytes("dram__bytes_read.sum"), "dram_write_bytes": nbytes("dram__bytes_write.sum"),
                      "duration_us_under_ncu": t_us, "capture": os.path.basename(rep), "launch": note}, indent=1))


def main() -> None:
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic":
        traffic_json(sys.argv[1], sys.argv[3])
        return
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    print(f"# {rep}")
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        print(f"kernel  {d['Kernel Name'][:120]}")
        print(f"grid {d['Grid Size']} block {d['Block Size']}")
        for k, label in KEYS:
            if d.get(k, "") != "":
                print(f"  {label:34s} {d[k]:>16s} {u[k]}")
        stalls = []
        for k in head:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                # both "warp_latency" and "warps" spellings exist across ncu versions
                try:
                    stalls.append((float(d[k]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
            elif k.startswith("smsp__average_warp_latency_issue_stalled_") or (
                    k.startswith("smsp__average_warps_issue_stalled") and k.endswith(".ratio")):
                try:
                    stalls.append((float(d[k]), k.split("issue_stalled_")[-1].replace(".ratio", "")))
                except ValueError:
                    pass
        seen = set()
        top = []
        for v, n in sorted(stalls, reverse=True):
            if n not in seen:
                seen.add(n)
                top.append(f"{n} {v:.2f}")
        if top:
            print("  top stalls (warps per issue)      " + "; ".join(top[:6]))
        print()


if __name__ == "__main__":
    main()
