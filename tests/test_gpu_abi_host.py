"""The "_host" entry points of the C ABI through RAW ctypes — the binding INTEGRATION.md §2 shows a
maintainer of the reference, not the package's own _lib / Engine wrappers: host buffers in, host
buffers out, nothing but the shared library between the test and the GPU."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "datmo_using_optical_flow_b200", "lib", "libdatmo_b200.so")


class FarnebackParams(C.Structure):
    _fields_ = [("pyr_scale", C.c_double), ("levels", C.c_int), ("winsize", C.c_int), ("iterations", C.c_int),
                ("poly_n", C.c_int), ("poly_sigma", C.c_double), ("flags", C.c_int), ("variant", C.c_int)]


@pytest.fixture(scope="module")
def raw():
    from datmo_using_optical_flow_b200 import build
    build.build()
    lib = C.CDLL(LIB)
    lib.datmo_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.datmo_destroy.argtypes = [C.c_void_p]
    lib.datmo_last_error.argtypes = [C.c_void_p]
    lib.datmo_last_error.restype = C.c_char_p
    lib.datmo_farneback_default_params.argtypes = [C.POINTER(FarnebackParams)]
    lib.datmo_farneback_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(FarnebackParams), C.c_void_p]
    lib.datmo_bev_bins.argtypes = [C.c_double, C.c_double, C.c_double]
    lib.datmo_bev_rasterize_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_double,
                                             C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                             C.c_double, C.c_void_p]
    handle = C.c_void_p()
    assert lib.datmo_create(0, None, C.byref(handle)) == 0
    yield lib, handle
    lib.datmo_destroy(handle)


def _farneback(lib, handle, prev, nxt, dtype):
    p = FarnebackParams()
    lib.datmo_farneback_default_params(C.byref(p))
    B, H, W = prev.shape
    flow = np.empty((B, H, W, 2), np.float32)
    st = lib.datmo_farneback_host(handle, prev.ctypes.data, nxt.ctypes.data, dtype, H, W, B, C.byref(p), flow.ctypes.data)
    assert st == 0, lib.datmo_last_error(handle).decode()
    return flow


def test_farneback_host_vs_cv2(raw):
    import cv2
    from datmo_using_optical_flow_b200 import synth
    lib, handle = raw
    pairs = [synth.textured_pair(s, 300, 420, shift=(1 + s, -2)) for s in range(3)]
    prev = np.ascontiguousarray(np.stack([a for a, _ in pairs]))
    nxt = np.ascontiguousarray(np.stack([b for _, b in pairs]))
    for _ in range(2):      # the second call reuses the handle's staging buffers
        flow = _farneback(lib, handle, prev, nxt, 0)
    for i in range(3):
        ref = cv2.calcOpticalFlowFarneback(prev[i].astype(np.float32), nxt[i].astype(np.float32), None,
                                           0.3, 5, 15, 5, 5, 5.0, 0)
        d = np.abs(flow[i] - ref)
        assert d.max() <= 1e-3 and d.mean() <= 1e-5, (i, d.max(), d.mean())
    # float32 frames take the same path and give the same flow
    flow_f = _farneback(lib, handle, prev.astype(np.float32), nxt.astype(np.float32), 1)
    assert np.array_equal(flow_f, flow)
    # a bad argument comes back as a status code and a message, not an abort
    p = FarnebackParams()
    lib.datmo_farneback_default_params(C.byref(p))
    p.flags = 256
    st = lib.datmo_farneback_host(handle, prev.ctypes.data, nxt.ctypes.data, 0, 300, 420, 3, C.byref(p), flow.ctypes.data)
    assert st == -1 and b"flags" in lib.datmo_last_error(handle)


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_bev_rasterize_host_vs_reference_golden(raw, golden, case):
    lib, handle = raw
    g = golden("bev.npz")
    pts = np.ascontiguousarray(g[f"{case}_points"], dtype=np.float64)
    rx, ry, x0, x1, y0, y1, hmax = (float(v) for v in g[f"{case}_params"])
    nx, ny = lib.datmo_bev_bins(x0, x1, rx), lib.datmo_bev_bins(y0, y1, ry)
    want = g[f"{case}_bev"]
    assert (nx, ny) == want.shape
    bev = np.zeros((nx, ny), np.uint8)
    st = lib.datmo_bev_rasterize_host(handle, pts.ctypes.data, 0, len(pts), rx, ry, x0, y0, nx, ny, 0.5, 0.5, hmax,
                                      bev.ctypes.data)
    assert st == 0, lib.datmo_last_error(handle).decode()
    assert np.array_equal(bev, want)


class ChainConfig(C.Structure):
    _fields_ = [("H", C.c_int), ("W", C.c_int), ("batch", C.c_int), ("dtype", C.c_int), ("px_x", C.c_double),
                ("px_y", C.c_double), ("alpha_cont", C.c_double), ("thresh", C.c_double), ("eps", C.c_double),
                ("min_samples", C.c_int), ("cap", C.c_int), ("max_clusters", C.c_int), ("want_cells", C.c_int),
                ("n_slots", C.c_int), ("fb", FarnebackParams)]


def test_flow_to_clusters_host_vs_reference_calls(raw):
    """datmo_flow_to_clusters_host, the one-call form of main.py:577-615, bound as INTEGRATION.md §2 writes it:
    moving cells vs the reference chain on cv2's flow (guard band for cells at a threshold), and labels
    identical to sklearn DBSCAN fed the filtered velocities the GPU itself produced (checked through the
    device-resident chain in test_gpu_stages; here: cluster count and the partition's sizes)."""
    from datmo_using_optical_flow_b200 import synth
    from oracle import reference_port
    lib, handle = raw
    lib.datmo_chain_default_config.argtypes = [C.POINTER(ChainConfig)]
    lib.datmo_flow_to_clusters_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ChainConfig), C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    B, H, W = 2, 240, 320
    prev, nxt = synth.bev_pairs(40, B, H, W)
    cfg = ChainConfig()
    lib.datmo_chain_default_config(C.byref(cfg))
    cfg.H, cfg.W, cfg.batch, cfg.px_x, cfg.px_y, cfg.cap, cfg.max_clusters = H, W, B, 0.1, 0.1, H * W, 512
    n_valid, n_clusters = np.zeros(B, np.int32), np.zeros(B, np.int32)
    offsets = np.zeros(B + 1, np.int64)
    labels = np.zeros(B * H * W, np.int32)
    indices = np.zeros((B * H * W, 2), np.int32)
    summary = np.zeros((B, 512, 8), np.float64)
    for _ in range(2):    # the second call reuses the chain cached on the handle
        st = lib.datmo_flow_to_clusters_host(handle, prev.ctypes.data, nxt.ctypes.data, C.byref(cfg), n_valid.ctypes.data,
                                             n_clusters.ctypes.data, offsets.ctypes.data, labels.ctypes.data,
                                             indices.ctypes.data, labels.size, summary.ctypes.data)
        assert st == 0, lib.datmo_last_error(handle).decode()
    xr, yr = [-0.05 * W, 0.05 * W], [-0.05 * H, 0.05 * H]
    for b in range(B):
        want = reference_port.flow_to_clusters(prev[b], nxt[b], xr, yr, 1.0, 0.2, 5.0, 3)
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        assert hi - lo == n_valid[b]
        got_idx = indices[lo:hi]
        sym = len(set(map(tuple, got_idx.tolist())) ^ set(map(tuple, want["indices"].tolist())))
        assert sym <= max(3, 0.002 * len(want["indices"])), (b, sym)
        assert abs(int(n_clusters[b]) - len(want["clusters"])) <= 1
        lab = labels[lo:hi]
        assert lab.max() + 1 == n_clusters[b]
        # the summary's per-cluster counts are the label histogram
        k = int(n_clusters[b])
        assert np.array_equal(summary[b, :k, 0], np.bincount(lab[lab >= 0], minlength=k))
        assert (summary[b, k:] == 0).all()
    # too small a caller array: status code, counts still written
    st = lib.datmo_flow_to_clusters_host(handle, prev.ctypes.data, nxt.ctypes.data, C.byref(cfg), n_valid.ctypes.data,
                                         n_clusters.ctypes.data, offsets.ctypes.data, labels.ctypes.data,
                                         indices.ctypes.data, 10, summary.ctypes.data)
    assert st == -3 and offsets[B] > 10
