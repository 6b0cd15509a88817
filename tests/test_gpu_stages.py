"""GPU parity of the non-Farneback stages, through the C ABI: BEV rasteriser (bit-exact), velocity /
masks (exact), DBSCAN (labels identical to sklearn's), cluster summaries, RANSAC scoring, fused
preprocessing."""
import numpy as np
import pytest
import torch

from datmo_using_optical_flow_b200 import main, synth
from oracle import bev_np, cluster_np, dbscan_np, masks_np, ransac_np, reference_port

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


# ---- BEV ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_bev_golden_bit_exact(engine, golden, case):
    g = golden("bev.npz")
    p = g[f"{case}_points"]
    rx, ry, x0, x1, y0, y1, hmax = (float(v) for v in g[f"{case}_params"])
    got = main.compute_bev_grid(p, [rx, ry], [x0, x1], [y0, y1], h_max=hmax, engine=engine)
    assert got.dtype == np.uint8 and np.array_equal(got, g[f"{case}_bev"])


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2"])
def test_bev_large_cloud_bit_exact_both_layouts(engine, cfg):
    c = synth.SWEEP_CONFIGS[cfg]
    pts = synth.lidar_sweep(0, 3, c["beams"], c["n_points"], c["n_movers"])
    rng = np.random.default_rng(1)
    p64 = np.repeat(pts[:, :3].astype(np.float64), 10, axis=0) + rng.normal(scale=0.01, size=(len(pts) * 10, 3))
    want = bev_np.compute_bev_grid(p64, c["grid_resolution"], c["x_range"], c["y_range"], h_max=2.0)
    got = main.compute_bev_grid(p64, c["grid_resolution"], c["x_range"], c["y_range"], h_max=2.0, engine=engine)
    assert got.shape == want.shape and np.array_equal(got, want), int((got != want).sum())
    # float32 xyzw layout: same values as the f64 view of the f32 points
    want32 = bev_np.compute_bev_grid(pts[:, :3].astype(np.float64), c["grid_resolution"], c["x_range"], c["y_range"], h_max=2.0)
    got32 = main.compute_bev_grid(dev(pts), c["grid_resolution"], c["x_range"], c["y_range"], h_max=2.0, engine=engine)
    assert np.array_equal(host(got32), want32)


def test_bev_empty_and_ragged(engine):
    got = main.compute_bev_grid(np.zeros((0, 3)), [0.5, 0.5], [-2, 2], [-1, 1], engine=engine)
    assert got.shape == (8, 4) and not got.any()
    one = np.array([[0.1, 0.1, 1.0]])
    got = main.compute_bev_grid(one, [0.5, 0.5], [-2, 2], [-1, 1], engine=engine)
    assert np.array_equal(got, bev_np.compute_bev_grid_loops(one, [0.5, 0.5], [-2, 2], [-1, 1]))


# ---- velocity / masks --------------------------------------------------------------------------------
def test_velocity_and_masks_exact(engine, golden):
    g = golden("flow_chain.npz")
    xr, yr = [float(v) for v in g["ranges"][:2]], [float(v) for v in g["ranges"][2:]]
    for name in ("tex", "blob"):
        vx, vy = g[f"{name}_vx"], g[f"{name}_vy"]
        H, W = vx.shape
        # feed the golden velocities back as a "flow" with unit pixel size: every derived grid must be exact
        m = main.continuity_mask(vx, vy, 0.2, engine=engine)
        assert m.dtype == np.int64 and np.array_equal(m.astype(np.uint8), g[f"{name}_mask"])
        vxf, vyf, mag, angf, valid = main.moving_cell_filter(vx, vy, 0.2, engine=engine)
        ovxf, ovyf, omag, oangf, ovalid = masks_np.moving_cell_filter(vx, vy, masks_np.continuity_mask(vx, vy, 0.2))
        assert vxf.dtype == np.float64 and np.array_equal(vxf, ovxf) and np.array_equal(vyf, ovyf)
        assert np.array_equal(valid, g[f"{name}_valid"])
        assert np.array_equal(mag, omag)
        assert np.array_equal(angf.astype(np.float32), g[f"{name}_angf"].astype(np.float32))
    # pixel-size scaling and curl from a raw flow field
    rng = np.random.default_rng(0)
    flow = rng.uniform(-4, 4, (2, 50, 70, 2)).astype(np.float32)
    vm = engine.velocity_mask(dev(flow), 0.25, 0.2, 0.2)
    for b in range(2):
        ovx, ovy, oang = masks_np.flow_to_velocity(flow[b], [0.0, 17.5], [0.0, 10.0])     # 17.5/70 = .25, 10/50 = .2
        assert np.array_equal(host(vm["vx"])[b], ovx) and np.array_equal(host(vm["vy"])[b], ovy)
        assert np.array_equal(host(vm["ang"])[b], oang)
        om = masks_np.continuity_mask(ovx, ovy, 0.2)
        assert np.array_equal(host(vm["mask"])[b], om.astype(np.uint8))
        valid = masks_np.moving_cell_filter(ovx, ovy, om)[4]
        assert np.array_equal(host(vm["valid"])[b].astype(bool), valid)
        assert int(host(vm["n_valid"])[b]) == int(valid.sum())


@pytest.mark.parametrize("k", range(4))
def test_propagation_masks_exact(engine, golden, k):
    """main.py:166-221 through the C ABI: identical to the reference's double loops (golden) and, on a
    1024^2 field, to the oracle; f32 and f64 inputs keep the reference's dtype semantics."""
    g = golden("propagation.npz")
    dt, gx, gy, alpha = (float(v) for v in g[f"params_{k}"])
    vx, vy, ax, ay = (g[f"{n}_{k}"] for n in ("vx", "vy", "ax", "ay"))
    m = main.propagation_mask(vx, vy, dt, [gx, gy], alpha, engine=engine)
    assert m.dtype == np.int64 and np.array_equal(m, g[f"mask_{k}"])
    m = main.propagation_mask_with_acceleration(vx, vy, ax, ay, dt, [gx, gy], alpha, engine=engine)
    assert np.array_equal(m, g[f"mask_acc_{k}"])


def test_propagation_mask_large_batched(engine):
    rng = np.random.default_rng(9)
    B, H, W = 3, 1024, 1024
    vx = (rng.normal(size=(B, H, W)) * 1.5).astype(np.float32)
    vy = (rng.normal(size=(B, H, W)) * 1.5).astype(np.float32)
    vx[:, 100:400, 200:700] = 0.8          # a coherent mover: many sources per target band
    vy[:, 100:400, 200:700] = -0.45
    vx[0, 5, 5], vy[0, 6, 6] = np.nan, np.inf    # the reference raises on these; here they do not propagate
    got = host(engine.propagation_mask(dev(vx), dev(vy), 1.0, [0.1, 0.1], 0.2))
    for b in range(B):
        want = masks_np.propagation_mask(vx[b], vy[b], 1.0, [0.1, 0.1], 0.2)
        assert np.array_equal(got[b].astype(np.int64), want)


# ---- DBSCAN ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("i", range(5))
def test_dbscan_golden_labels_identical(engine, golden, i):
    g = golden("dbscan.npz")
    eps, ms = g[f"params_{i}"]
    labels, idx = main.dbscan_clustering(g["vx"].astype(np.float64), g["vy"].astype(np.float64), g["valid"],
                                         eps=float(eps), min_samples=int(ms), engine=engine)
    assert labels.dtype == np.intp and idx.dtype == np.int64
    assert np.array_equal(idx, g["indices"])
    assert np.array_equal(labels, g[f"labels_{i}"])
    assert dbscan_np.same_partition(labels, g[f"labels_{i}"])


def test_dbscan_random_fields_vs_sklearn_batched(engine):
    rng = np.random.default_rng(7)
    B, H, W = 6, 90, 130
    vx = np.where(rng.uniform(size=(B, H, W)) < 0.2, rng.uniform(-2, 2, (B, H, W)), 0).astype(np.float32)
    vy = np.where(vx != 0, rng.uniform(-2, 2, (B, H, W)), 0).astype(np.float32)
    valid = np.sqrt(vx.astype(np.float64) ** 2 + vy.astype(np.float64) ** 2) > 0.1
    for eps, ms in [(5.0, 3), (1.0, 2), (2.5, 6), (1.5, 4), (7.5, 5), (0.5, 1)]:   # 7.5: the rolled far-window path; 0.5: r = 0
        n_valid, labels, indices, n_clusters = engine.dbscan_grid(dev(vx), dev(vy), dev(valid.astype(np.uint8)), eps, ms)
        n_valid, labels, indices, n_clusters = (host(t) for t in (n_valid, labels, indices, n_clusters))
        for b in range(B):
            want, widx = dbscan_np.dbscan_clustering_sklearn(vx[b].astype(np.float64), vy[b].astype(np.float64), valid[b], eps, ms)
            n = n_valid[b]
            assert n == len(want)
            assert np.array_equal(indices[b, :n], widx)
            assert np.array_equal(labels[b, :n], want), (eps, ms, b)
            assert n_clusters[b] == (want.max() + 1 if len(want) and want.max() >= 0 else 0)


def _dbscan_fields(rng, B, H, W, kind):
    if kind == "blobs":       # rectangles with holes, smooth velocities: long runs, few roots
        valid = np.zeros((B, H, W), bool)
        for b in range(B):
            for _ in range(10):
                y, x = rng.integers(0, H - 5), rng.integers(0, W - 5)
                valid[b, y:y + rng.integers(2, 14), x:x + rng.integers(2, 40)] = True
        valid &= rng.uniform(size=(B, H, W)) < 0.97
        vx = rng.normal(size=(B, H, W)) * 0.4
        vy = rng.normal(size=(B, H, W)) * 0.4
    elif kind == "ties":      # integer velocities: d2 lands exactly on eps^2 (dr = eps, dv = 0)
        valid = rng.uniform(size=(B, H, W)) < 0.5
        vx = np.round(rng.normal(size=(B, H, W)) * 2.0)
        vy = np.round(rng.normal(size=(B, H, W)) * 2.0)
    else:                     # dense noise with large jumps: broken links inside rows
        valid = rng.uniform(size=(B, H, W)) < 0.8
        vx = rng.normal(size=(B, H, W)) * 3.0
        vy = rng.normal(size=(B, H, W)) * 3.0
    return (vx * valid).astype(np.float32), (vy * valid).astype(np.float32), valid


@pytest.mark.parametrize("impl", ["runs", "cells"])
@pytest.mark.parametrize("kind", ["blobs", "ties", "jumps"])
def test_dbscan_run_rule_vs_sklearn(engine, kind, impl, monkeypatch):
    """Both implementations (row runs on bit planes; cell-level passes) against live sklearn on fields
    that stress the run rule: long runs, exact ties at eps, broken links; widths that are not a
    multiple of 32 and wider than one 8-word segment."""
    if impl == "cells":
        monkeypatch.setenv("DATMO_DBSCAN_CELLS", "1")
    rng = np.random.default_rng({"blobs": 11, "ties": 12, "jumps": 13}[kind])
    B, H, W = 4, 61, 300
    vx, vy, valid = _dbscan_fields(rng, B, H, W, kind)
    for eps, ms in [(5.0, 3), (1.0, 2), (3.3, 4), (2.0, 5), (15.9, 9)]:
        n_valid, labels, indices, n_clusters = engine.dbscan_grid(dev(vx), dev(vy), dev(valid.astype(np.uint8)), eps, ms)
        n_valid, labels, indices, n_clusters = (host(t) for t in (n_valid, labels, indices, n_clusters))
        for b in range(B):
            want, widx = dbscan_np.dbscan_clustering_sklearn(vx[b].astype(np.float64), vy[b].astype(np.float64),
                                                             valid[b], eps, ms)
            n = n_valid[b]
            assert n == len(want)
            assert np.array_equal(indices[b, :n], widx)
            assert np.array_equal(labels[b, :n], want), (kind, impl, eps, ms, b)
            assert n_clusters[b] == (want.max() + 1 if want.max() >= 0 else 0)


def test_dbscan_bench_density_1024_labels_identical_to_sklearn(engine):
    """The benchmarked regime: a 1024x1024 synth.bev_pair through the GPU chain (150-260 k moving cells, ~140
    clusters); the GPU's own vx_f / vy_f / valid fed to the reference's sklearn call must give the same labels."""
    a, b = synth.bev_pair(1, 1024, 1024)
    res = engine.flow_pipeline(dev(a)[None], dev(b)[None], 0.1, 0.1, 0.2, 5.0, 3, max_clusters=0)
    n = int(host(res.n_valid)[0])
    assert n >= 100_000
    vx_f, vy_f, valid = host(res.vx_f)[0], host(res.vy_f)[0], host(res.valid)[0].astype(bool)
    want, widx = dbscan_np.dbscan_clustering_sklearn(vx_f.astype(np.float64), vy_f.astype(np.float64), valid, 5.0, 3)
    assert n == len(want)
    assert np.array_equal(host(res.indices)[0, :n], widx)
    assert np.array_equal(host(res.labels)[0, :n], want)
    assert int(host(res.n_clusters)[0]) == want.max() + 1


def test_dbscan_large_grid_vs_grid_rule_oracle(engine):
    """cfg3-sized grid, sparse moving cells; the oracle grid rule is itself pinned to sklearn on CPU."""
    rng = np.random.default_rng(3)
    H = W = 1024
    vx = np.zeros((H, W), np.float32)
    vy = np.zeros((H, W), np.float32)
    for _ in range(60):
        y, x = rng.integers(0, H - 40), rng.integers(0, W - 40)
        h, w = rng.integers(3, 30), rng.integers(3, 30)
        vx[y:y + h, x:x + w] = rng.uniform(-1, 1) + rng.uniform(-0.05, 0.05, (h, w))
        vy[y:y + h, x:x + w] = rng.uniform(-1, 1) + rng.uniform(-0.05, 0.05, (h, w))
    valid = np.sqrt(vx.astype(np.float64) ** 2 + vy.astype(np.float64) ** 2) > 0.1
    want, widx = dbscan_np.dbscan_grid(vx, vy, valid, 5.0, 3)
    labels, idx = main.dbscan_clustering(vx, vy, valid, eps=5.0, min_samples=3, engine=engine)
    assert np.array_equal(idx, widx) and np.array_equal(labels, want)


def test_dbscan_empty_mask_raises_like_sklearn(engine):
    z = np.zeros((16, 16), np.float32)
    with pytest.raises(ValueError):
        main.dbscan_clustering(z, z, np.zeros((16, 16), bool), engine=engine)


def test_dbscan_capacity_truncates_but_counts(engine):
    vx = np.ones((8, 8), np.float32)
    n_valid, labels, indices, _ = engine.dbscan_grid(dev(vx), dev(vx), dev(np.ones((8, 8), np.uint8)), 1.0, 2, cap=10)
    assert int(host(n_valid)[0]) == 64 and labels.shape == (1, 10)
    assert (host(labels)[0] == 0).all()


# ---- clusters ---------------------------------------------------------------------------------------
def test_cluster_summary_vs_oracle(engine, golden):
    g = golden("flow_chain.npz")
    mask = g["blob_mask"].astype(np.int64)
    vx_f, vy_f = g["blob_vx"] * mask, g["blob_vy"] * mask
    got = main.extract_cluster_data(g["blob_labels"], g["blob_indices"], vx_f, vy_f, engine=engine)
    want = cluster_np.extract_cluster_data(g["blob_labels"], g["blob_indices"], vx_f, vy_f)
    assert sorted(got) == sorted(want) == list(g["blob_cl_keys"])
    for k in want:
        np.testing.assert_allclose(got[k]["measurement"], want[k]["measurement"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(np.sort(got[k]["eigenvalues"]), np.sort(want[k]["eigenvalues"]), rtol=1e-7, atol=1e-7,
                                   equal_nan=True)
    with pytest.raises(ValueError):
        main.extract_cluster_data(g["blob_labels"][:-1], g["blob_indices"], vx_f, vy_f, engine=engine)


def test_flow_to_clusters_chain_vs_reference_port(engine):
    a, b = synth.bev_pair(12, 240, 320)
    xr, yr = [-16.0, 16.0], [-12.0, 12.0]
    labels, indices, clusters = main.flow_to_clusters(a, b, xr, yr, 1.0, 0.2, 5.0, 3, engine=engine)
    want = reference_port.flow_to_clusters(a, b, xr, yr, 1.0, 0.2, 5.0, 3)
    # flows agree to ~1e-4, so cells within that of a threshold may flip: compare with a guard band
    sym = len(set(map(tuple, indices.tolist())) ^ set(map(tuple, want["indices"].tolist())))
    assert sym <= max(3, 0.002 * len(want["indices"])), (sym, len(want["indices"]))
    assert abs(len(clusters) - len(want["clusters"])) <= 1


# ---- RANSAC ------------------------------------------------------------------------------------------
def test_ransac_hypotheses_and_scores_vs_oracle(engine):
    pts = synth.lidar_sweep(1, 0, 32, 20_000, 1)
    n = len(pts)
    iters = 700
    out = engine.ransac_ground(dev(pts), 0.5, 5, iters, seed=11, flip_x=True, return_hypotheses=True)
    planes, cnt, err = host(out["hyp_planes"]), host(out["hyp_count"]), host(out["hyp_err"])
    p64 = pts[:, :3].astype(np.float64)
    p64[:, 0] = -p64[:, 0]
    idx, ok = ransac_np.sample_indices(11, iters, 5, n)
    assert ok.all()
    want_planes = ransac_np.plane_from_points(p64[idx])
    np.testing.assert_allclose(planes, want_planes, rtol=0, atol=1e-9)
    # scoring is exact given the device's own planes
    wc, we = ransac_np.score_planes(p64, planes, 0.5)
    assert np.array_equal(cnt, wc)
    np.testing.assert_allclose(err, we, rtol=1e-10, atol=1e-9)
    best = host(out["best"])
    assert best[0] == ransac_np.select_best(wc, we) and best[1] == wc[best[0]]
    assert np.array_equal(host(out["plane"]), planes[best[0]])
    mask = host(out["inlier_mask"]).astype(bool)
    assert np.array_equal(mask, ransac_np.point_plane_distance(p64, planes[best[0]]) < 0.5)
    refit = host(out["refit"])
    want_refit = ransac_np.plane_from_points(p64[mask][None])[0]
    np.testing.assert_allclose(refit * np.sign(refit[2]), want_refit * np.sign(want_refit[2]), atol=1e-7)
    # scene-level: the planted ground plane z = -2.5 is found
    nrm = refit[:3] * np.sign(refit[2])
    assert np.degrees(np.arccos(np.clip(nrm[2], -1, 1))) < 0.5 and abs(abs(refit[3]) - 2.5) < 0.05


@pytest.mark.parametrize("flip", [False, True])
def test_ransac_f32_bounds_pass_picks_the_exact_winner(engine, monkeypatch, flip):
    """The production path brackets every hypothesis' inlier count in f32 and re-scores only the possible
    winners in fp64; winner, count, plane, refit and inlier mask must be those of the full fp64 pass
    (which the hypothesis-level test above pins to the numpy oracle)."""
    for seq, n_pts in ((1, 20_000), (4, 120_000)):
        pts = dev(synth.lidar_sweep(seq, 0, 32, n_pts, 2))
        fast = engine.ransac_ground(pts, 0.5, 5, 5000, seed=11, flip_x=flip)
        monkeypatch.setenv("DATMO_RANSAC_EXACT", "1")
        full = engine.ransac_ground(pts, 0.5, 5, 5000, seed=11, flip_x=flip)
        monkeypatch.delenv("DATMO_RANSAC_EXACT")
        # the refit's moments are summed without atomics, in a fixed order: bit-equal too
        for k in ("best", "plane", "inlier_mask", "refit"):
            assert np.array_equal(host(fast[k]), host(full[k])), (seq, k)
        again = engine.ransac_ground(pts, 0.5, 5, 5000, seed=11, flip_x=flip)
        assert np.array_equal(host(again["refit"]), host(fast["refit"]))      # and run-to-run reproducible
        assert host(fast["best"])[1] > 0.3 * len(pts)        # the ground plane won


def test_ransac_f64_layout_and_three_point_model(engine):
    rng = np.random.default_rng(2)
    p = np.column_stack([rng.uniform(-20, 20, 5000), rng.uniform(-20, 20, 5000), 1.0 + rng.normal(0, 0.01, 5000)])
    p[:1000, 2] = rng.uniform(2, 8, 1000)
    out = engine.ransac_ground(dev(p), 0.3, 3, 200, seed=5, return_hypotheses=True)
    planes = host(out["hyp_planes"])
    idx, ok = ransac_np.sample_indices(5, 200, 3, len(p))
    np.testing.assert_allclose(planes, ransac_np.plane_from_points(p[idx]), atol=1e-9)
    wc, _ = ransac_np.score_planes(p, planes, 0.3)
    assert np.array_equal(host(out["hyp_count"]), wc)
    assert host(out["inlier_mask"])[1000:].mean() > 0.99


# ---- fused preprocessing --------------------------------------------------------------------------------
def test_preprocess_fused_vs_oracle_given_noise_and_ground(engine):
    c = synth.SWEEP_CONFIGS["cfg1"]
    pts = synth.lidar_sweep(2, 1, c["beams"], c["n_points"], c["n_movers"])
    n = len(pts)
    rng = np.random.default_rng(4)
    noise = rng.normal(scale=0.01, size=(n, 10, 3))
    ground = np.abs(pts[:, 2] + 2.5) < 0.5
    roi = [-50, 50, -50, 50, -3, 1]
    got = main.preprocess_points(pts, c["grid_resolution"], c["x_range"], c["y_range"], 2.0, roi, noise=noise,
                                 ground_mask=ground, engine=engine)
    p64 = pts[:, :3].astype(np.float64)
    p64[:, 0] = -p64[:, 0]
    keep = (~ground) & (p64[:, 0] >= -50) & (p64[:, 0] <= 50) & (p64[:, 1] >= -50) & (p64[:, 1] <= 50) & \
           (p64[:, 2] >= -3) & (p64[:, 2] <= 1)
    want = reference_port.preprocess_points(pts, c["grid_resolution"], c["x_range"], c["y_range"], 2.0, roi,
                                            noise[keep].reshape(-1, 3), ground_mask=ground)
    assert got.shape == want.shape and np.array_equal(got, want), int((got != want).sum())


def test_preprocess_with_device_ransac_and_device_noise(engine):
    c = synth.SWEEP_CONFIGS["cfg1"]
    pts = synth.lidar_sweep(2, 2, c["beams"], c["n_points"], c["n_movers"])
    roi = [-50, 50, -50, 50, -3, 1]
    bev = main.preprocess_points(pts, c["grid_resolution"], c["x_range"], c["y_range"], 2.0, roi, seed=3, engine=engine)
    assert bev is not None and bev.shape == (400, 400) and bev.dtype == np.uint8 and bev.max() == 255
    occ = (bev > 0).mean()
    assert 0.001 < occ < 0.2        # ground removed: only objects remain
    # same seed -> same hypotheses, same noise, and accumulators that do not depend on the order in which the
    # warps arrive (128-bit fixed-point sums): the grid is reproducible bit for bit
    for _ in range(3):
        bev2 = main.preprocess_points(pts, c["grid_resolution"], c["x_range"], c["y_range"], 2.0, roi, seed=3, engine=engine)
        assert np.array_equal(bev, bev2)
    # empty ROI -> None, like the reference (main.py:84-86)
    assert main.preprocess_points(pts, c["grid_resolution"], c["x_range"], c["y_range"], 2.0,
                                  [1000, 1001, 1000, 1001, 0, 1], engine=engine) is None


# ---- ROI crop / density expansion as device ops ---------------------------------------------------
def test_roi_filter_and_expansion_device(engine, golden):
    g = golden("bev.npz")
    got = main.filter_points_in_roi(g["roi_points"], [float(v) for v in g["roi_bounds"]], engine=engine)
    assert np.array_equal(got, g["roi_out"])
    rng = np.random.default_rng(8)
    p = rng.uniform(-12, 12, (50_000, 3))
    roi = [-10, 10, -10, 10, -3, 1]
    assert np.array_equal(main.filter_points_in_roi(p, roi, engine=engine), bev_np.filter_points_in_roi(p, roi))
    p4 = np.concatenate([p, rng.uniform(size=(len(p), 1))], axis=1).astype(np.float32)
    keep = ((p4[:, 0] >= -10) & (p4[:, 0] <= 10) & (p4[:, 1] >= -10) & (p4[:, 1] <= 10) & (p4[:, 2] >= -3) & (p4[:, 2] <= 1))
    assert np.array_equal(host(engine.roi_filter(dev(p4), roi)), p4[keep])
    assert main.filter_points_in_roi(np.zeros((0, 3)), roi, engine=engine).shape == (0, 3)
    noise = rng.normal(scale=0.01, size=(len(p) * 10, 3))
    exp = main.increase_point_density(p, 10, 0.01, noise=noise, engine=engine)
    assert np.array_equal(exp, bev_np.increase_point_density(p, 10, noise=noise))
    drawn = main.increase_point_density(p[:2000], 10, 0.01, seed=5, engine=engine)
    d = drawn - np.repeat(p[:2000], 10, axis=0)
    assert abs(d.std() - 0.01) < 5e-4 and abs(d.mean()) < 3e-4          # N(0, 0.01) on every coordinate


# ---- the sequence driver (preprocess -> flow -> clusters -> tracker) ---------------------------------
def test_sequence_pipeline_tracks_the_mover(engine, tmp_path):
    from datmo_using_optical_flow_b200.pipeline import process_clouds
    cfg = dict(grid_resolution=[0.25, 0.25], x_range=[-50.0, 50.0], y_range=[-50.0, 50.0], z_max=2.0,
               roi_bounds=[-50, 50, -50, 50, -3, 1], dt=1.0)
    clouds = [synth.lidar_sweep(3, f, 32, 60_000, 1, dt=0.1) for f in range(4)]
    out = process_clouds(clouds, cfg, engine=engine, seed=1, output_dir=str(tmp_path), save_grids=True)
    assert len(out["bevs"]) == 4 and all(b is not None and b.shape == (400, 400) for b in out["bevs"])
    assert len(out["pairs"]) == 3
    done = [p for p in out["pairs"] if not p["skipped"]]
    assert done, "every pair was skipped"
    for p in done:
        assert len(p["labels"]) == len(p["indices"]) and p["tracks"].shape[1] == 6
        assert set(p["clusters"]) == set(range(len(p["clusters"])))
    # the reference's per-frame files (saving_utils.py names and dtypes)
    for f in range(4):
        assert np.array_equal(np.load(tmp_path / f"bev_frame_{f}.npy"), out["bevs"][f])
    for p in done:
        k = p["index"]
        assert np.array_equal(np.load(tmp_path / f"dbscan_labels_frame_{k}.npy"), p["labels"])
        assert np.array_equal(np.load(tmp_path / f"dbscan_indices_frame_{k}.npy"), p["indices"])
        vx = np.load(tmp_path / f"velocity_x_frame_{k}.npy")
        assert vx.dtype == np.float64 and vx.shape == (400, 400)
        assert (tmp_path / f"ekf_tracks_frame_{k}.yaml").exists()
    assert (tmp_path / "tracks.csv").read_text().startswith("Frame Index,Track ID,Linear Velocity")
    assert (tmp_path / "filtered_velocities.csv").read_text().startswith("Frame Index,Point Index,Filtered X Velocity")


def test_process_sequences_batched_equals_single_sequence_runs(engine):
    """BASELINE configs[4] in miniature: three sequences advanced in lock step — one batched flow -> clusters call
    per tick — must give, for every sequence, exactly the tracks of driving that sequence on its own
    (process_clouds): sharding / batching must not change results."""
    from datmo_using_optical_flow_b200.pipeline import process_clouds, process_sequences
    cfg = dict(grid_resolution=[0.25, 0.25], x_range=[-50.0, 50.0], y_range=[-50.0, 50.0], z_max=2.0,
               roi_bounds=[-50, 50, -50, 50, -3, 1], dt=1.0)
    n_seq, n_frames = 3, 5
    seqs = [[synth.lidar_sweep(10 + s, f, 32, 30_000, 1 + s, dt=0.1) for f in range(n_frames)] for s in range(n_seq)]
    seqs[1][2] = None            # a dropped frame: pairs 1 and 2 of that sequence are skipped (main.py:572-574)
    out = process_sequences(seqs, cfg, engine=engine, seed=4)
    assert out["local"] == [0, 1, 2] and len(out["ticks"]) == n_frames and len(out["gathered"]) == n_frames
    n_done = 0
    for s in range(n_seq):
        single = process_clouds([c for c in seqs[s]], cfg, engine=engine, seed=4 + 1000 * s)
        for k in range(1, n_frames):
            rec, want = out["ticks"][k][s], single["pairs"][k - 1]
            assert rec["skipped"] == want["skipped"], (s, k, rec.get("error"))
            if not rec["skipped"]:
                n_done += 1
                assert np.array_equal(rec["tracks"], want["tracks"]) and np.array_equal(rec["saved_tracks"], want["saved_tracks"])
                assert sorted(rec["clusters"]) == sorted(want["clusters"])
        assert np.array_equal(out["gathered"][-1][s], single["tracks"].as_array())
    assert n_done >= 6
    assert out["runner"].last_ms["pairs"] >= 2


@pytest.mark.parametrize("want_cells", [True, False])
def test_host_flow_pipeline_matches_device_chain(engine, want_cells):
    """The host-buffer entry (bench.py's e2e leg, a thin caller of the C-ABI chain): pinned uint8 in, compact
    labels (int16) / packed cell indices / summaries out — or counts and summaries only — against the
    device-resident chain, three double-buffered batches."""
    from datmo_using_optical_flow_b200.engine import HostFlowPipeline, farneback_params
    B, H, W = 3, 200, 240
    pipe = HostFlowPipeline(engine, B, H, W, 0.25, 0.25, 0.2, 5.0, 3, farneback_params(), cap=H * W, max_clusters=256,
                            want_cells=want_cells)
    batches = [synth.bev_pairs(10 * k, B, H, W) for k in range(3)]
    pins = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in batches]
    outs = []
    pipe.submit(0, *pins[0])
    with pytest.raises(Exception):
        pipe.submit(0, *pins[0])          # the slot is in flight
    for k in range(3):
        if k + 1 < 3:
            pipe.submit((k + 1) % 2, *pins[k + 1])
        nv, ncl, off, lab, idx, summ = pipe.collect(k % 2)
        outs.append((nv.copy(), ncl.copy(), off.copy(), lab.copy(), HostFlowPipeline.unpack_indices(idx).copy(), summ.copy()))
    assert pipe.d2h_bytes > 0 and not pipe.truncated
    for k, (a, b) in enumerate(batches):
        res = engine.flow_pipeline(dev(a), dev(b), 0.25, 0.25, 0.2, 5.0, 3, farneback_params(), cap=H * W,
                                   max_clusters=256)
        nv, ncl, off, lab, idx, summ = outs[k]
        assert np.array_equal(nv, host(res.n_valid)) and np.array_equal(ncl, host(res.n_clusters))
        assert np.array_equal(np.diff(off), nv)
        if want_cells:
            assert lab.dtype == np.int16
            for i in range(B):
                n = int(nv[i])
                assert np.array_equal(lab[off[i]:off[i] + n], host(res.labels)[i, :n])
                assert np.array_equal(idx[off[i]:off[i] + n], host(res.indices)[i, :n])
        else:
            assert len(lab) == 0
        kmax = summ.shape[1]
        assert kmax == int(ncl.max())
        assert np.array_equal(summ, host(res.summary)[:, :kmax], equal_nan=True)
    pipe.close()
