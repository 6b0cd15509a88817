"""Sequence driver: the reference's ``process_multiple_frames`` loop
(/root/reference/Optical_flow/main.py:541-641) with the hot path on the GPU.

Differences from the reference loop, all deliberate:
  * lines 588-589 of main.py (the unguarded acceleration that raises on the first pair
    and makes every pair fail) are not reproduced; ax / ay were dead values anyway;
  * each frame is preprocessed once and reused as bev2 of pair i-1 and bev1 of pair i
    (the reference preprocesses it twice with different expansion noise);
  * the per-pair results are returned; with ``output_dir`` the data files the reference saves
    (.npy grids, DBSCAN arrays, track YAML, the two CSVs; artefacts.py) are written too — the
    matplotlib PNGs are out of scope.
Per-pair failures are caught and the pair skipped, like the reference's try/except.
"""
from __future__ import annotations

from . import artefacts
from . import main as ops
from .engine import Engine, default_engine
from .tracker import TrackManager

# the reference's config.yaml keys that the loop reads (main.py:542-550)
DEFAULT_CONFIG = dict(grid_resolution=[0.2, 0.2], x_range=[-20, 20], y_range=[-20, 20], z_max=2.0,
                      roi_bounds=[-10, 10, -10, 10, -3, 1], masks=dict(alpha_p=[0.8], alpha_cont=[0.2]), dt=1.0,
                      dbscan_params=dict(eps=5.0, min_samples=3))


def process_clouds(clouds, config=None, engine=None, seed=0, ground_masks=None, verbose=False, output_dir=None,
                   save_grids=False):
    """clouds: iterable of float32 (N,4) sweeps of ONE sequence (or .pcd paths).
    Returns dict(tracks=TrackManager, pairs=[per-pair dict], bevs=[uint8 grids or None]).
    output_dir: also write the reference's per-frame files there (saving_utils.py formats);
    save_grids adds the filtered velocity grids and the per-cell CSV (large)."""
    cfg = dict(DEFAULT_CONFIG)
    cfg.update(config or {})
    eng = engine or default_engine()
    alpha_cont = cfg["masks"]["alpha_cont"][0]
    dbp = cfg["dbscan_params"]
    tm = TrackManager()
    bevs, pairs = [], []
    prev = None
    if output_dir is not None:
        import os
        os.makedirs(output_dir, exist_ok=True)
        tracks_csv = os.path.join(output_dir, "tracks.csv")
        cells_csv = os.path.join(output_dir, "filtered_velocities.csv")
        for f in (tracks_csv, cells_csv):
            if os.path.exists(f):
                os.remove(f)   # the reference starts its CSV afresh (main.py:556-558)
    for i, cloud in enumerate(clouds):
        try:
            pts = ops.read_pcd(cloud) if isinstance(cloud, (str, bytes)) else cloud
            gm = None if ground_masks is None else ground_masks[i]
            bev = ops.preprocess_points(pts, cfg["grid_resolution"], cfg["x_range"], cfg["y_range"], cfg["z_max"],
                                        cfg["roi_bounds"], seed=seed + i, ground_mask=gm, engine=eng)
        except Exception as exc:  # the reference prints and moves on (main.py:635-637)
            if verbose:
                print(f"Error preprocessing frame {i}: {exc}")
            bev = None
        bevs.append(bev)
        if output_dir is not None and bev is not None:
            artefacts.save_bev(output_dir, bev, i)
        if i == 0:
            prev = bev
            continue
        rec = dict(index=i - 1, skipped=True)
        if prev is None or bev is None:
            if verbose:
                print(f"Invalid BEV grid for frames {i - 1} and {i}. Skipping.")
        else:
            try:
                out = ops.flow_to_clusters(prev, bev, cfg["x_range"], cfg["y_range"], cfg["dt"], alpha_cont,
                                           dbp["eps"], dbp["min_samples"], engine=eng,
                                           return_grids=output_dir is not None and save_grids)
                labels, indices, clusters = out[:3]
                if len(labels) == 0:
                    raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by DBSCAN.")
                # main.py:618-634: association + EKF, THEN the savers, THEN lifetimes and manage_tracks — a
                # confirmed track that manage_tracks deletes on this pair is still in this pair's files
                tm.associate_and_update(clusters, cfg["dt"])
                saved_tracks = tm.as_array()
                if output_dir is not None:
                    k = i - 1
                    if save_grids:
                        # main.py:600-609: the filtered field, its magnitude and curl (all from the device)
                        g = out[3]
                        artefacts.save_velocity_grid(output_dir, g["vx_filtered"], g["vy_filtered"], k)
                        artefacts.save_all_filtered_velocities_to_csv(g["vx_filtered"], g["vy_filtered"],
                                                                      g["velocity_magnitude"], g["angular_velocity"],
                                                                      k, cells_csv)
                    artefacts.save_dbscan_results(output_dir, labels, indices, k)
                    artefacts.save_ekf_tracks(output_dir, tm, k)
                    artefacts.save_all_velocities_to_csv(tm, k, tracks_csv)
                tm.step_lifetimes()
                rec = dict(index=i - 1, skipped=False, labels=labels, indices=indices, clusters=clusters,
                           tracks=tm.as_array(), saved_tracks=saved_tracks)
            except Exception as exc:
                if verbose:
                    print(f"Error processing frames {i - 1} and {i}: {exc}")
        pairs.append(rec)
        prev = bev
    return dict(tracks=tm, pairs=pairs, bevs=bevs)


# ---------------------------------------------------------------------------------------------------
# several sequences at once (BASELINE configs[4]: 8 concurrent 128-beam sequences at 20 Hz)
# ---------------------------------------------------------------------------------------------------
class SequenceRunner:
    """The reference's driver loop (main.py:541-641) for SEVERAL sequences advancing in lock step on one
    GPU.  Sequences are independent (the EKF state belongs to one sequence), so on every tick the k-th
    sweep of each sequence is preprocessed to a device-resident BEV and the pairs (previous BEV, this BEV)
    of ALL sequences go through flow -> velocity -> mask -> DBSCAN -> cluster summaries as ONE batched call;
    only the per-cluster summaries (a few KB) come back to the host, where each sequence's tracker
    (tracker.TrackManager, the reference's EKF semantics) consumes them.  One process per GPU owns a
    shard of the sequences (sharding.shard_sequences); `sharding.gather_sequence_tracks` collects the
    tracks of all shards once per tick.

    sequence_ids: the job-wide ids of the sequences this runner owns (they seed the expansion noise, so a
    sequence gives the same result however the job is sharded).  tick(clouds) takes one float32 (N,4)
    sweep (numpy, or a CUDA tensor) per owned sequence — None for a dropped frame — and returns one record
    per sequence."""

    def __init__(self, sequence_ids, config=None, engine=None, seed: int = 0, max_clusters: int = 512,
                 cap: int | None = None, keep_cells: bool = False, ground_masks=None, pre_streams: int | None = None):
        import torch
        self.torch = torch
        self.cfg = dict(DEFAULT_CONFIG)
        self.cfg.update(config or {})
        self.eng = engine or default_engine()
        self.ids = [int(s) for s in sequence_ids]
        self.n = len(self.ids)
        self.seed = int(seed)
        self.max_clusters = int(max_clusters)
        self.keep_cells = bool(keep_cells)
        self.ground_masks = ground_masks     # optional {sequence id: [per-frame uint8 masks]} instead of RANSAC
        self.trackers = [TrackManager() for _ in range(self.n)]
        # The sweeps of one tick are independent: their cloud -> BEV chains (a dozen short, latency-bound kernels
        # around one long one, per sweep) run on a few engines of their own — own handle, stream and workspace —
        # so that they overlap on the GPU; the batched flow -> clusters call then waits for all of them.
        # datmo_preprocess_dev blocks its caller until the sweep's point count is known (an empty ROI is a return
        # code), so each engine is driven by a host thread of its own; ctypes drops the GIL for the call.
        k = min(4, self.n) if pre_streams is None else int(pre_streams)
        self.pre_engines = [Engine(self.eng.device, isolated=True) for _ in range(k)] if k > 1 else []
        self._pool = None
        if self.pre_engines:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=k, thread_name_prefix="datmo-pre")
        self.prev = [None] * self.n          # device BEV of the previous tick, per sequence
        self.frame = 0
        eng, cfg = self.eng, self.cfg
        self.nx = eng.bev_bins(cfg["x_range"][0], cfg["x_range"][1], cfg["grid_resolution"][0])
        self.ny = eng.bev_bins(cfg["y_range"][0], cfg["y_range"][1], cfg["grid_resolution"][1])
        self.cap = self.nx * self.ny if cap is None else int(cap)
        # main.py:147-150: pixel size = range / shape (the x range over axis 1, as the reference does)
        self.px = (cfg["x_range"][1] - cfg["x_range"][0]) / self.ny
        self.py = (cfg["y_range"][1] - cfg["y_range"][0]) / self.nx
        self.last_ms = {}

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        for pe in self.pre_engines:
            pe.close()
        self.pre_engines = []

    def _bev(self, j: int, cloud):
        torch, cfg = self.torch, self.cfg
        eng = self.pre_engines[j % len(self.pre_engines)] if self.pre_engines else self.eng
        if cloud is None:
            return None
        try:
            with eng.on_stream():       # the upload too runs on the engine's stream (an isolated engine orders nothing)
                pts = cloud if isinstance(cloud, torch.Tensor) else torch.from_numpy(cloud)
                pts = pts.to(eng.tdev, dtype=torch.float32, non_blocking=True)
                if pts.shape[1] == 3:
                    pts = torch.cat([pts, torch.zeros_like(pts[:, :1])], dim=1)
                gm = None
                if self.ground_masks is not None:
                    gm = torch.as_tensor(self.ground_masks[self.ids[j]][self.frame]).to(eng.tdev)
                # the same seed process_clouds(seed=seed + 1000 * id) gives frame `self.frame` of this sequence
                return eng.preprocess(pts, cfg["grid_resolution"], cfg["x_range"], cfg["y_range"], cfg["z_max"],
                                      cfg["roi_bounds"], seed=self.seed + 1000 * self.ids[j] + self.frame,
                                      ground_mask=gm)
        except Exception:       # the reference prints and moves on (main.py:635-637)
            return None

    def tick(self, clouds):
        import time
        import numpy as np
        torch, eng, cfg = self.torch, self.eng, self.cfg
        if len(clouds) != self.n:
            raise ValueError(f"expected {self.n} sweeps, one per sequence")
        t0 = time.perf_counter()
        if self._pool is not None:
            k = len(self.pre_engines)

            def group(e):       # one thread per engine: the sweeps e, e + k, .. in order
                with torch.cuda.device(self.eng.device):
                    return [(j, self._bev(j, clouds[j])) for j in range(e, self.n, k)]

            bevs = [None] * self.n
            for part in self._pool.map(group, range(k)):
                for j, b in part:
                    bevs[j] = b
        else:
            bevs = [self._bev(j, c) for j, c in enumerate(clouds)]
        for pe in self.pre_engines:     # host-side join: the BEVs are complete before the batched call reads them
            pe.synchronize()
        eng.synchronize()
        t1 = time.perf_counter()
        live = [j for j in range(self.n) if self.prev[j] is not None and bevs[j] is not None]
        recs = [dict(sequence=self.ids[j], frame=self.frame, skipped=True) for j in range(self.n)]
        t2 = t3 = t1
        if live:
            with eng.on_stream():
                prev = torch.stack([self.prev[j] for j in live])
                cur = torch.stack([bevs[j] for j in live])
            res = eng.flow_pipeline(prev, cur, self.px, self.py, cfg["masks"]["alpha_cont"][0], cfg["dbscan_params"]["eps"],
                                    cfg["dbscan_params"]["min_samples"], cap=self.cap, max_clusters=self.max_clusters,
                                    keep_flow=False)
            eng.synchronize()
            ncl = res.n_clusters.cpu().numpy()
            nval = res.n_valid.cpu().numpy()
            kmax = int(min(ncl.max(), self.max_clusters))
            summ = res.summary[:, :kmax].cpu().numpy() if kmax else np.zeros((len(live), 0, 8))
            t2 = time.perf_counter()
            for k, j in enumerate(live):
                rec = recs[j]
                try:
                    if nval[k] == 0:
                        raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by DBSCAN.")
                    clusters = ops.clusters_from_summary(summ[k], int(ncl[k]), self.max_clusters)
                    tm = self.trackers[j]
                    tm.associate_and_update(clusters, cfg["dt"])       # main.py:618
                    saved = tm.as_array()                                # what the savers see (main.py:619-620)
                    tm.step_lifetimes()                                  # main.py:621-634
                    rec.update(skipped=False, n_valid=int(nval[k]), n_clusters=int(ncl[k]), clusters=clusters,
                               saved_tracks=saved, tracks=tm.as_array())
                    if self.keep_cells:
                        n = int(min(nval[k], self.cap))
                        rec.update(labels=res.labels[k, :n].cpu().numpy().astype(np.intp),
                                   indices=res.indices[k, :n].cpu().numpy().astype(np.int64))
                except Exception as exc:       # per-pair failures skip the pair (main.py:635-637)
                    rec["error"] = f"{type(exc).__name__}: {exc}"
            t3 = time.perf_counter()
        for j in range(self.n):
            self.prev[j] = bevs[j]
        self.frame += 1
        self.last_ms = dict(preprocess=1e3 * (t1 - t0), flow_to_summaries=1e3 * (t2 - t1), tracker=1e3 * (t3 - t2),
                            total=1e3 * (t3 - t0), pairs=len(live))
        return recs


def process_sequences(sequences, config=None, engine=None, seed=0, max_clusters=512, gather=True, max_tracks=64,
                      ground_masks=None):
    """sequences: list (over ALL sequences of the job) of equally long lists of float32 (N,4) sweeps.
    Under torch.distributed each rank takes its shard (sharding.shard_sequences), runs its sequences in
    lock step on its GPU and, once per tick, all-gathers the track tables (NCCL; the only collective).
    Returns dict(local=[sequence ids], ticks=[per tick: one record per local sequence],
    gathered=[per tick: {sequence id: (n,6) track table}], identical on every rank)."""
    import torch.distributed as dist
    from . import sharding
    on = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    local = sharding.shard_sequences(len(sequences), rank, world)
    runner = SequenceRunner(local, config, engine, seed=seed, max_clusters=max_clusters, ground_masks=ground_masks)
    n_ticks = len(sequences[0]) if sequences else 0
    ticks, gathered = [], []
    try:
        for k in range(n_ticks):
            recs = runner.tick([sequences[s][k] for s in local])
            ticks.append(recs)
            if gather:
                tables = {s: runner.trackers[j].as_array() for j, s in enumerate(local)}
                gathered.append(sharding.gather_sequence_tracks(tables, len(sequences), max_tracks))
    finally:
        runner.close()      # the helper engines and their host threads; trackers and records stay usable
    return dict(local=local, ticks=ticks, gathered=gathered, runner=runner)
