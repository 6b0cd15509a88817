"""CPU: host-side logic — tracker vs the reference's EKF golden, sharding + gloo gathers
(world_size 2), PCD reader, synthetic generators."""
import os
import socket
import struct
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from datmo_using_optical_flow_b200 import sharding, synth
from datmo_using_optical_flow_b200.tracker import TrackManager

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- tracker ------------------------------------------------------------------------------------
def test_tracker_matches_reference_golden(golden):
    g = golden("tracks.npz")
    tm = TrackManager()
    deleted_now = 0
    for f in range(int(g["n_frames"])):
        cen, meas, eig = g[f"f{f}_centroid"], g[f"f{f}_meas"], g[f"f{f}_eig"]
        clusters = {k: dict(centroid=cen[k], measurement=meas[k].tolist(), eigenvalues=eig[k]) for k in range(len(cen))}
        # the two halves of the reference's loop body (main.py:618 | 621-634) with the savers in between
        tm.associate_and_update(clusters, 1.0)
        saved = g[f"f{f}_saved"]      # what save_ekf_tracks / the tracks CSV hold for this pair (main.py:619-620)
        got_saved = np.array([[tid, *t.state.tolist()] for tid, t in tm.tracks.items()], dtype=np.float64).reshape(-1, 5)
        assert got_saved.shape == saved.shape, f
        np.testing.assert_allclose(got_saved, saved, rtol=1e-12, atol=1e-12, err_msg=f"saved, frame {f}")
        deleted_now += len(saved) - len(g[f"f{f}_tracks"])
        tm.step_lifetimes()
        want = g[f"f{f}_tracks"]
        got = tm.as_array()
        assert got.shape == want.shape, f
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, err_msg=f"frame {f}")
    assert len(tm.confirmed) >= 1
    # the stream is long enough for manage_tracks to delete a confirmed track (lifetime 16..25): on that pair
    # the saved table still holds it, the post-manage table does not
    assert int(g["n_frames"]) >= 17 and deleted_now >= 1


def test_tracker_reference_quirks():
    tm = TrackManager()
    far = {0: dict(centroid=np.array([10.0, 10.0]), measurement=[10.0, 10.0, 0.1, 0.0], eigenvalues=np.array([3.0, 3.0])),
           1: dict(centroid=np.array([50.0, 50.0]), measurement=[50.0, 50.0, 0.0, 0.1], eigenvalues=np.array([3.0, 3.0]))}
    tm.update(far, 1.0)
    # both unmatched clusters get id 1 (max of the OLD table + 1): only the last survives
    assert list(tm.tracks) == [1] and tm.tracks[1].state[0] == 50.0
    tm.update({}, 1.0)
    assert tm.tracks == {} and tm.lifetimes == {}          # unmatched tracks are dropped
    nan = {0: dict(centroid=np.array([1.0, 1.0]), measurement=[1.0, 1.0, 0, 0], eigenvalues=np.array([np.nan, np.nan]))}
    tm.update(nan, 1.0)
    tm.update(nan, 1.0)                                       # NaN distance never matches -> always a new track
    assert list(tm.tracks) == [2]
    assert tm.as_array(max_tracks=3).shape == (3, 6) and np.isnan(tm.as_array(max_tracks=3)[1:]).all()


# ---- sharding -----------------------------------------------------------------------------------
@pytest.mark.parametrize("total,world", [(4096, 8), (4096, 3), (8, 8), (5, 8), (0, 2), (100, 1)])
def test_shard_range_partitions(total, world):
    got = []
    sizes = []
    for r in range(world):
        s, c = sharding.shard_range(total, r, world)
        got += list(range(s, s + c))
        sizes.append(c)
    assert got == list(range(total))
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_shard_sequences_cfg5():
    assert [sharding.shard_sequences(8, r, 4) for r in range(4)] == [[0, 1], [2, 3], [4, 5], [6, 7]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank "processes" its shard of 11 pairs and of 3 sequences
        s, c = sharding.shard_range(11, rank, world)
        local_sum = float(sum(range(s, s + c)))
        red = sharding.reduce_metrics({"pairs": c, "checksum": local_sum, "t": 1.0 + rank}, "sum")
        mx = sharding.reduce_metrics({"t": 1.0 + rank}, "max")
        tracks = np.array([[10 * rank + i, rank, i, 0.5, -0.5, i % 2] for i in range(rank + 2)], dtype=np.float64)
        gathered = sharding.gather_tracks(tracks, max_tracks=4)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), pairs=red["pairs"], checksum=red["checksum"], tmax=mx["t"],
                 n=len(gathered), **{f"g{i}": g for i, g in enumerate(gathered)})
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_reduce(tmp_path):
    world = 2
    mp.spawn(_gloo_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        assert d["pairs"] == 11 and d["checksum"] == sum(range(11)) and d["tmax"] == 2.0
        assert d["n"] == 2
        assert d["g0"].shape == (2, 6) and d["g1"].shape == (3, 6)      # unpadded, in rank order, on every rank
        assert d["g1"][2, 0] == 12.0


def _gloo_sequences_worker(rank, world, port, out_dir):
    """The N > 1 host logic of pipeline.process_sequences without a GPU: every rank drives the trackers of its
    shard of 5 sequences with canned cluster dictionaries and all-gathers the track tables once per tick."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_seq, n_ticks = 5, 6
        local = sharding.shard_sequences(n_seq, rank, world)
        tms = {s: TrackManager() for s in local}
        out = {}
        for k in range(n_ticks):
            for s in local:
                tms[s].update(_canned_clusters(s, k), 1.0)
            got = sharding.gather_sequence_tracks({s: tms[s].as_array() for s in local}, n_seq, max_tracks=8)
            assert sorted(got) == list(range(n_seq))
            for s, t in got.items():
                out[f"t{k}_s{s}"] = t
        np.savez(os.path.join(out_dir, f"seq_r{rank}.npz"), **out)
    finally:
        dist.destroy_process_group()


def _canned_clusters(seq, tick):
    rng = np.random.default_rng(100 * seq + 7)
    base = rng.uniform(20, 80, (3, 2))
    cl = {}
    for j in range(1 + (seq + tick) % 3):
        c = base[j] + 0.01 * tick
        cl[j] = dict(centroid=c, measurement=[c[0], c[1], 0.0, 0.0], eigenvalues=np.array([0.02, 0.03]))
    return cl


def test_gloo_world2_sequence_shards_gather_every_tick(tmp_path):
    world = 2
    mp.spawn(_gloo_sequences_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "seq_r0.npz"), np.load(tmp_path / "seq_r1.npz")
    assert sorted(r0.files) == sorted(r1.files) and len(r0.files) == 5 * 6
    # the single-process run of every sequence
    for s in range(5):
        tm = TrackManager()
        for k in range(6):
            tm.update(_canned_clusters(s, k), 1.0)
            want = tm.as_array()
            for r in (r0, r1):       # every rank holds every sequence's table, equal to the unsharded run
                assert np.array_equal(r[f"t{k}_s{s}"], want)


def test_gather_sequence_tracks_single_process():
    tabs = {0: np.arange(12, dtype=np.float64).reshape(2, 6), 1: np.zeros((0, 6))}
    got = sharding.gather_sequence_tracks(tabs, 2, max_tracks=4)
    assert np.array_equal(got[0], tabs[0]) and got[1].shape == (0, 6)


def test_clusters_from_summary_reference_behaviours():
    from datmo_using_optical_flow_b200.main import clusters_from_summary
    s = np.array([[5, 10.0, 20.0, 0.5, -0.25, 2.0, 0.5, 1.0], [3, 1.0, 2.0, 0.0, 0.0, 1.0, 0.0, 1.0]])
    cl = clusters_from_summary(s, 2, 8)
    assert list(cl) == [0, 1] and cl[0]["measurement"] == [10.0, 20.0, 0.5, -0.25]
    lam = np.linalg.eigvals(np.array([[2.0, 0.5], [0.5, 1.0]]))
    assert np.allclose(np.sort(cl[0]["eigenvalues"]), np.sort(lam))
    one = np.array([[1, 3.0, 4.0, 0.1, 0.1, np.nan, np.nan, np.nan]])
    with pytest.raises(np.linalg.LinAlgError):      # the reference's eigvals raises on the NaN covariance (main.py:424)
        clusters_from_summary(one, 1, 8)
    with pytest.raises(RuntimeError):               # more clusters than summary rows is an error, not a truncation
        clusters_from_summary(s, 9, 8)


def test_gather_tracks_single_process_and_truncation():
    t = np.arange(30, dtype=np.float64).reshape(5, 6)
    out = sharding.gather_tracks(t, max_tracks=3)
    assert len(out) == 1 and np.array_equal(out[0], t[:3])
    assert sharding.gather_tracks(np.zeros((0, 6)), 4)[0].shape == (0, 6)


# ---- PCD reader ------------------------------------------------------------------------------------
def _write_pcd(path, pts, binary):
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
           f"WIDTH {len(pts)}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {len(pts)}\nDATA {'binary' if binary else 'ascii'}\n")
    with open(path, "wb") as fh:
        fh.write(hdr.encode())
        if binary:
            fh.write(pts.astype("<f4").tobytes())
        else:
            for p in pts:
                fh.write(("%.9g %.9g %.9g\n" % tuple(p)).encode())


@pytest.mark.parametrize("binary", [False, True])
def test_read_pcd_roundtrip(tmp_path, binary):
    from datmo_using_optical_flow_b200.main import read_pcd
    pts = np.random.default_rng(0).uniform(-50, 50, (257, 3)).astype(np.float32)
    path = str(tmp_path / "frame.pcd")
    _write_pcd(path, pts, binary)
    got = read_pcd(path)
    assert got.dtype == np.float32 and got.shape == (257, 4)
    assert np.array_equal(got[:, :3], pts) and not got[:, 3].any()
    with open(path, "wb") as fh:
        fh.write(b"FIELDS a b\nDATA ascii\n1 2\n")
    with pytest.raises(ValueError):
        read_pcd(path)


# ---- synthetic inputs --------------------------------------------------------------------------------
def test_synth_is_deterministic_and_sized():
    a1, b1 = synth.bev_pair(7, 256, 320)
    a2, b2 = synth.bev_pair(7, 256, 320)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2) and a1.dtype == np.uint8 and a1.shape == (256, 320)
    assert (a1 != b1).any()
    p = synth.lidar_sweep(0, 0, 32, 60_000, 1)
    assert p.dtype == np.float32 and p.shape[1] == 4 and 40_000 < len(p) < 80_000
    assert np.array_equal(p, synth.lidar_sweep(0, 0, 32, 60_000, 1))
    ground = np.abs(p[:, 2] + 2.5) < 0.1
    assert 0.3 < ground.mean() < 0.98
    q = synth.lidar_sweep(0, 5, 32, 60_000, 1)
    assert len(q) != len(p) or not np.array_equal(p, q)          # the mover moved


# ---- artefact writers (saving_utils.py formats) ---------------------------------------------
def _fake_tracks():
    from datmo_using_optical_flow_b200.tracker import Track
    return {3: Track(state=np.array([10.5, 20.25, 0.125, -0.5])), 7: Track(state=np.array([1.0, 2.0, 3.0, 4.0]))}


def test_artefact_writers_formats(tmp_path):
    import csv
    import yaml
    from datmo_using_optical_flow_b200 import artefacts
    out = str(tmp_path)
    bev = (np.arange(12, dtype=np.uint8)).reshape(3, 4)
    artefacts.save_bev(out, bev, 5)
    assert np.array_equal(np.load(tmp_path / "bev_frame_5.npy"), bev)
    labels, idx = np.array([0, 0, -1], dtype=np.intp), np.array([[1, 2], [1, 3], [2, 0]], dtype=np.int64)
    artefacts.save_dbscan_results(out, labels, idx, 5)
    assert np.array_equal(np.load(tmp_path / "dbscan_labels_frame_5.npy"), labels)
    assert np.array_equal(np.load(tmp_path / "dbscan_indices_frame_5.npy"), idx)
    tracks = _fake_tracks()
    artefacts.save_ekf_tracks(out, tracks, 5)
    assert yaml.safe_load(open(tmp_path / "ekf_tracks_frame_5.yaml")) == {3: [10.5, 20.25, 0.125, -0.5], 7: [1.0, 2.0, 3.0, 4.0]}
    csv_file = str(tmp_path / "tracks.csv")
    artefacts.save_all_velocities_to_csv(tracks, 5, csv_file)
    artefacts.save_all_velocities_to_csv(tracks, 6, csv_file)      # appends, header once
    rows = list(csv.reader(open(csv_file)))
    assert rows[0] == artefacts.TRACK_CSV_HEADER and len(rows) == 5
    assert rows[1][:2] == ["5", "3"] and float(rows[1][2]) == np.linalg.norm([0.125, -0.5])
    assert [float(v) for v in rows[1][3:]] == [0.125, -0.5, 20.25]   # x vel, y vel, state[1] ("angular")
    vx = np.array([[0.0, 0.5], [0.0, 0.0]])
    vy = np.array([[0.0, 0.0], [-1.0, 0.0]])
    cells = str(tmp_path / "cells.csv")
    artefacts.save_all_filtered_velocities_to_csv(vx, vy, np.hypot(vx, vy), vx * 0 + 2.0, 9, cells)
    rows = list(csv.reader(open(cells)))
    assert rows[0] == artefacts.CELL_CSV_HEADER
    assert rows[1] == ["9", "0", "0.5", "0.0", "0.5", "2.0"] and rows[2] == ["9", "1", "0.0", "-1.0", "1.0", "2.0"]


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).reference_available(),
                    reason="reference tree not mounted")
def test_artefact_writers_match_reference_savers(tmp_path, monkeypatch):
    """Byte-for-byte against the reference's own CSV / YAML writers (saving_utils.py:80-103, 17-46, 119-125)."""
    import sys
    from datmo_using_optical_flow_b200 import artefacts
    from oracle import ref_loader
    ref = ref_loader.load_reference_main()
    su = sys.modules["saving_utils"]
    tracks = _fake_tracks()
    ours, theirs = tmp_path / "ours", tmp_path / "theirs"
    ours.mkdir(), theirs.mkdir()
    su.save_all_velocities_to_csv(tracks, 4, str(theirs / "t.csv"))
    artefacts.save_all_velocities_to_csv(tracks, 4, str(ours / "t.csv"))
    assert open(ours / "t.csv").read() == open(theirs / "t.csv").read()
    rng = np.random.default_rng(2)
    vx = rng.normal(size=(6, 7)) * (rng.uniform(size=(6, 7)) < 0.4)
    vy = rng.normal(size=(6, 7)) * (vx != 0)
    su.save_all_filtered_velocities_to_csv(vx, vy, np.hypot(vx, vy), vx - vy, 4, str(theirs / "c.csv"))
    artefacts.save_all_filtered_velocities_to_csv(vx, vy, np.hypot(vx, vy), vx - vy, 4, str(ours / "c.csv"))
    assert open(ours / "c.csv").read() == open(theirs / "c.csv").read()
    # save_ekf_tracks: the YAML half (the plotting half needs matplotlib, stubbed to no-ops here)
    plt = sys.modules["matplotlib.pyplot"]
    for name in ("figure", "plot", "quiver", "title", "xlabel", "ylabel", "legend", "grid", "savefig", "close"):
        monkeypatch.setattr(plt, name, lambda *a, **k: None, raising=False)
    monkeypatch.setattr(su, "output_dir", str(theirs))
    su.save_ekf_tracks(tracks, 4)
    artefacts.save_ekf_tracks(str(ours), tracks, 4)
    assert open(ours / "ekf_tracks_frame_4.yaml").read() == open(theirs / "ekf_tracks_frame_4.yaml").read()


def test_numa_binding_reads_sysfs_and_never_raises(tmp_path, monkeypatch):
    """bind_to_device_numa_node: the GPU's PCI address -> numa_node -> cpulist -> sched_setaffinity; every
    missing piece degrades to a no-op with a reason."""
    import os
    import types

    import torch

    from datmo_using_optical_flow_b200 import sharding

    assert sharding._parse_cpulist("0-2,5,7-8\n") == {0, 1, 2, 5, 7, 8}
    props = types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1B, pci_device_id=0)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: props)
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    allowed = sorted(os.sched_getaffinity(0))
    (node / "cpulist").write_text(f"{allowed[0]}\n")
    called = {}
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: called.update(pid=pid, cpus=set(cpus)))
    out = sharding.bind_to_device_numa_node(0, sysfs=str(tmp_path))
    assert out == {"node": 1, "cpus": 1} and called == {"pid": 0, "cpus": {allowed[0]}}
    (dev / "numa_node").write_text("-1\n")
    assert sharding.bind_to_device_numa_node(0, sysfs=str(tmp_path))["node"] is None
    assert sharding.bind_to_device_numa_node(0, sysfs=str(tmp_path / "nowhere"))["node"] is None
