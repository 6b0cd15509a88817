"""CPU: the numpy Farneback restatement against live cv2 (the library the reference calls at
main.py:142) and against golden flows produced through the reference's compute_velocity_vectors.

Tolerance (BASELINE.json north_star): max |dflow| <= 1e-3 px, mean <= 1e-5 px.  On textured
(well-conditioned) frames the restatement sits at ~3e-6 max.  On sparse blob / BEV-like frames the
problem itself is ill-conditioned where the window sees no texture: cv2 run twice with inputs that
differ by one f32 ulp (3e-5 on a 0..255 scale) moves its own flow by up to 1.3e-3 px, so for those
frames the gate is mean <= 1e-5, 99.9th percentile <= 1e-3 and max <= 1e-2."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from datmo_using_optical_flow_b200 import synth  # noqa: E402
from oracle import farneback_np as fb  # noqa: E402

REF = dict(pyr_scale=0.3, levels=5, winsize=15, iterations=5, poly_n=5, poly_sigma=5.0, flags=0)


def _cv(a, b, **p):
    cv2.setNumThreads(1)
    return cv2.calcOpticalFlowFarneback(a.astype(np.float32), b.astype(np.float32), None, **p)


def test_level_plan_matches_survey():
    sizes = lambda H, W, s, l: [(d["h"], d["w"]) for d in fb.level_plan(H, W, s, l)]
    assert sizes(200, 200, 0.3, 5) == [(60, 60), (200, 200)]
    assert sizes(400, 400, 0.3, 5) == [(36, 36), (120, 120), (400, 400)]
    assert sizes(1024, 1024, 0.3, 5) == [(92, 92), (307, 307), (1024, 1024)]
    assert sizes(2048, 2048, 0.3, 5) == [(55, 55), (184, 184), (614, 614), (2048, 2048)]
    assert sizes(800, 800, 0.5, 5) == [(50, 50), (100, 100), (200, 200), (400, 400), (800, 800)]
    assert [d["ksize"] for d in fb.level_plan(1024, 1024, 0.3, 5)] == [25, 7, 3]


@pytest.mark.parametrize("ks,s", [(3, 0.0), (7, 1.1666666666666667), (25, 5.055555555555555), (91, 18.02)])
def test_gaussian_kernel_bit_exact(ks, s):
    assert np.array_equal(fb.gaussian_kernel(ks, s), cv2.getGaussianKernel(ks, s, cv2.CV_32F).ravel())


def test_pyramid_primitives_close_to_cv2():
    a, _ = synth.bev_pair(1, 200, 240)
    f = a.astype(np.float32)
    for L in fb.level_plan(200, 240, 0.3, 5):
        g = cv2.GaussianBlur(f, (L["ksize"], L["ksize"]), L["sigma"], sigmaY=L["sigma"])
        assert np.abs(g - fb.gaussian_blur(f, L["ksize"], L["sigma"])).max() < 2e-4
        r = cv2.resize(g, (L["w"], L["h"]), interpolation=cv2.INTER_LINEAR)
        assert np.abs(r - fb.resize_linear(g, L["h"], L["w"])).max() < 1e-4


@pytest.mark.parametrize("H,W,kw", [
    (96, 128, {}),
    (123, 257, {}),
    (200, 200, dict(pyr_scale=0.5, levels=3, winsize=14, iterations=3, poly_n=7, poly_sigma=1.5)),
    (160, 160, dict(winsize=16, iterations=1, levels=1)),
])
def test_textured_frames_within_tolerance(H, W, kw):
    a, b = synth.textured_pair(H * 7 + W, H, W)
    p = dict(REF, **kw)
    d = np.abs(fb.calc_optical_flow_farneback(a, b, **p) - _cv(a, b, **p))
    assert d.max() <= 1e-3 and d.mean() <= 1e-5
    assert d.max() <= 1e-4          # what the restatement actually achieves


@pytest.mark.parametrize("H,W,seed", [(200, 200, 1), (400, 400, 2)])
def test_blob_frames_within_conditioning_floor(H, W, seed):
    a, b = synth.bev_pair(seed, H, W)
    d = np.abs(fb.calc_optical_flow_farneback(a, b, **REF) - _cv(a, b, **REF)).max(axis=2)
    assert d.mean() <= 1e-5
    assert np.quantile(d, 0.999) <= 1e-3
    assert d.max() <= 1e-2


def test_cv2_own_sensitivity_documents_the_floor():
    """cv2 against itself with a 1-ulp input perturbation: the max-error floor on blob frames."""
    a, b = synth.bev_pair(2, 400, 400)
    rng = np.random.default_rng(0)
    a32 = a.astype(np.float32)
    pert = a32 + rng.uniform(-3e-5, 3e-5, a32.shape).astype(np.float32)
    d = np.abs(_cv(a32, b, **REF) - _cv(pert, b, **REF))
    assert d.mean() < 1e-5
    assert d.max() > 1e-5           # ulp-level input noise is amplified well beyond ulp level


def test_oracle_within_north_star_tolerance_where_cv2_is_stable():
    """The tolerance statement of the GPU tests, applied to the numpy restatement: max 1e-3 / mean 1e-5 px on
    the pixels where cv2 is stable against a 1-ulp input perturbation; elsewhere cv2 itself moves by more."""
    from oracle import flow_stability
    a, b = synth.bev_pair(3, 800, 800)
    r = flow_stability.compare(fb.calc_optical_flow_farneback(a, b, **REF), a, b, REF)
    assert r["stable_fraction"] >= 0.75 and r["max_stable"] <= 1e-3 and r["mean_stable"] <= 1e-5, r
    assert r["max_all"] > 1e-3 and r["ref_self_max"] > 1e-3       # this frame does have unstable pixels
    assert r["max_unstable"] <= 20 * r["ref_self_max"], r


def test_identical_and_zero_frames():
    z = np.zeros((64, 80), np.uint8)
    assert not fb.calc_optical_flow_farneback(z, z, **REF).any()      # all-zero frames -> exactly zero
    a, _ = synth.textured_pair(4, 64, 80)
    f = fb.calc_optical_flow_farneback(a, a, **REF)
    ref = _cv(a, a, **REF)
    assert np.abs(f - ref).max() <= 1e-4
    assert np.abs(ref).max() > 1e-3   # the last row/column quirk: identical frames give non-zero flow


def test_golden_flow_chain(golden):
    g = golden("flow_chain.npz")
    xr, yr = g["ranges"][:2], g["ranges"][2:]
    for name, tol_max in (("tex", 1e-4), ("blob", 1e-2)):
        a, b = g[f"{name}_a"], g[f"{name}_b"]
        flow = fb.calc_optical_flow_farneback(a, b, **REF)
        H, W = a.shape
        vx = flow[..., 0] * np.float32((xr[1] - xr[0]) / W)
        vy = flow[..., 1] * np.float32((yr[1] - yr[0]) / H)
        d = np.maximum(np.abs(vx - g[f"{name}_vx"]), np.abs(vy - g[f"{name}_vy"]))
        assert d.mean() <= 1e-5 and d.max() <= tol_max
