"""Multi-GPU plumbing: one process per GPU, work sharded by frame pair or by sequence,
no collective on the data path; torch.distributed only gathers tracks and metrics.

Frame pairs are independent through BEV -> flow -> clusters; the EKF carries state
across the pairs of ONE sequence (/root/reference/Optical_flow/main.py:553-634), so
throughput workloads shard by pair and end-to-end workloads shard by sequence
(SURVEY.md §8 e).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of `total` items owned by `rank`: (start, count).  Blocks differ by at
    most one item and concatenate, in rank order, to range(total)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_sequences(n_sequences: int, rank: int, world: int) -> list[int]:
    start, count = shard_range(n_sequences, rank, world)
    return list(range(start, start + count))


def _parse_cpulist(text: str) -> set[int]:
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the format of /sys/devices/system/node/node*/cpulist)."""
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_device_numa_node(device_index: int, sysfs: str = "/sys") -> dict:
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off, so that the pinned staging
    buffers it allocates next (first touch) and the copy threads sit on the socket whose PCIe root the GPU
    uses: with one rank per GPU on a two-socket host, half the ranks otherwise DMA across the socket link.
    Returns what was done ({"node": n, "cpus": k} or {"node": None, "why": ...}); never raises."""
    import os

    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"{sysfs}/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read())
        if node < 0:
            return {"node": None, "why": "the platform reports no NUMA node for the GPU"}
        with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as fh:
            cpus = _parse_cpulist(fh.read()) & os.sched_getaffinity(0)
        if not cpus:
            return {"node": None, "why": f"no allowed CPU on node {node}"}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception as exc:  # missing sysfs entries, old torch without the pci_* properties, ...
        return {"node": None, "why": f"{type(exc).__name__}: {exc}"}


def _is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def gather_tracks(local: np.ndarray | torch.Tensor, max_tracks: int, device=None) -> list[np.ndarray]:
    """All-gather of per-shard track tables.  local: (n, 6) rows [id, s0, s1, s2, s3, confirmed].
    Every rank contributes a fixed (max_tracks, 6) float64 block (NaN padded) plus its count;
    returns the unpadded table of every rank, in rank order, on every rank."""
    t = torch.as_tensor(np.asarray(local, dtype=np.float64).reshape(-1, 6))
    n = min(len(t), max_tracks)
    buf = torch.full((max_tracks * 6 + 1,), float("nan"), dtype=torch.float64)
    buf[0] = n
    buf[1:1 + n * 6] = t[:n].reshape(-1)
    if not _is_dist():
        return [t[:n].numpy()]
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    buf = buf.to(dev)
    out = [torch.empty_like(buf) for _ in range(dist.get_world_size())]
    dist.all_gather(out, buf)
    res = []
    for o in out:
        o = o.cpu()
        k = int(o[0].item())
        res.append(o[1:1 + k * 6].reshape(k, 6).numpy())
    return res


def gather_sequence_tracks(local: dict, n_sequences: int, max_tracks: int, device=None) -> dict:
    """One all-gather per tick for every sequence of the job.  local: {sequence id: (n, 6) track table} of
    the sequences this rank owns (a contiguous shard, sharding.shard_sequences).  Every rank contributes a
    fixed block of ceil(n_sequences / world) slots x (2 + max_tracks * 6) float64 (sequence id, row count,
    rows; NaN padded) and receives {sequence id: table} for all sequences."""
    world = dist.get_world_size() if _is_dist() else 1
    slots = -(-n_sequences // world) if n_sequences else 0
    width = 2 + max_tracks * 6
    buf = torch.full((max(slots, 1), width), float("nan"), dtype=torch.float64)
    buf[:, 0] = -1
    for k, (sid, table) in enumerate(sorted(local.items())):
        t = torch.as_tensor(np.asarray(table, dtype=np.float64).reshape(-1, 6))
        n = min(len(t), max_tracks)
        buf[k, 0], buf[k, 1] = sid, n
        buf[k, 2:2 + n * 6] = t[:n].reshape(-1)
    if not _is_dist():
        blocks = [buf]
    else:
        dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                                 if dist.get_backend() == "nccl" else torch.device("cpu"))
        send = buf.to(dev)
        out = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(out, send)
        blocks = [o.cpu() for o in out]
    res = {}
    for blk in blocks:
        for row in blk:
            sid = int(row[0].item())
            if sid < 0:
                continue
            n = int(row[1].item())
            res[sid] = row[2:2 + n * 6].reshape(n, 6).numpy().copy()
    return res


def reduce_metrics(values: dict[str, float], op: str = "sum", device=None) -> dict[str, float]:
    """All-reduce a small dict of scalars (timing, parity counters)."""
    keys = sorted(values)
    t = torch.tensor([float(values[k]) for k in keys], dtype=torch.float64)
    if _is_dist():
        dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                                 if dist.get_backend() == "nccl" else torch.device("cpu"))
        t = t.to(dev)
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op])
        t = t.cpu()
    return {k: float(v) for k, v in zip(keys, t.tolist())}
