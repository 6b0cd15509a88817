"""Device-side engine: one libdatmo_b200 handle bound to one GPU and one stream.

torch is plumbing here — device memory, streams, pinned host buffers — the
compute is the CUDA library behind the C ABI (include/datmo_b200.h).  All
methods take and return CUDA tensors; the numpy-facing functions with the
reference's names live in ``datmo_using_optical_flow_b200.main``.
"""
from __future__ import annotations

import ctypes as C
import math
from contextlib import contextmanager
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

# Optical_flow/main.py:132-140 — the values the reference hard-codes
REFERENCE_FARNEBACK = dict(pyr_scale=0.3, levels=5, winsize=15, iterations=5, poly_n=5, poly_sigma=5.0, flags=0)


def farneback_params(**kw) -> _lib.FarnebackParams:
    p = _lib.FarnebackParams()
    _lib.load().datmo_farneback_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown Farneback parameter {k!r}")
        setattr(p, k, v)
    return p


def farneback_layers_host(H: int, W: int, params: dict | None = None):
    """[(h, w)] of the pyramid layers the library processes, coarsest first (datmo_farneback_layers;
    host-side arithmetic only, no device needed)."""
    p = farneback_params(**{k: v for k, v in (params or {}).items()})
    w = (C.c_int * 16)()
    h = (C.c_int * 16)()
    n = _lib.load().datmo_farneback_layers(H, W, C.byref(p), 16, w, h)
    return [(h[i], w[i]) for i in range(n)]


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


@dataclass
class FlowPipelineResult:
    """Device-resident results of one batch of frame pairs."""
    flow: torch.Tensor | None        # [B,H,W,2] f32
    vx_f: torch.Tensor               # [B,H,W] f32, velocity * continuity mask
    vy_f: torch.Tensor
    valid: torch.Tensor              # [B,H,W] u8
    n_valid: torch.Tensor            # [B] i32
    labels: torch.Tensor             # [B,cap] i32
    indices: torch.Tensor            # [B,cap,2] i32 (row, col)
    n_clusters: torch.Tensor         # [B] i32
    summary: torch.Tensor | None     # [B,max_clusters,8] f64
    cap: int
    ang_f: torch.Tensor | None = None   # [B,H,W] f32 curl of the filtered field (main.py:604-606), on request


class Engine:
    """One per (device, stream); not thread-safe (same rule as the C handle)."""

    def __init__(self, device: int | None = None, isolated: bool = False):
        """isolated=True: the engine's stream is NOT ordered against the caller's current stream
        (no wait_stream on entry / exit), so several engines can run concurrently on one GPU; the
        caller then orders inputs / outputs itself (events or a synchronize)."""
        self.isolated = isolated
        if not torch.cuda.is_available():
            raise _lib.DatmoLibraryError("no CUDA device: datmo_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tdev = torch.device("cuda", self.device)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream(device=self.device)
        h = C.c_void_p()
        st = self.lib.datmo_create(self.device, C.c_void_p(self.stream.cuda_stream), C.byref(h))
        if st != _lib.OK:
            raise _lib.DatmoError(st, "datmo_create failed (is this an sm_100 device?)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.datmo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing -----------------------------------------------------------------------
    def _check(self, st: int):
        if st != _lib.OK:
            raise _lib.DatmoError(st, self.lib.datmo_last_error(self.h).decode(errors="replace"))

    @contextmanager
    def on_stream(self):
        """Run torch allocations / copies and library launches on the engine's stream,
        ordered after the caller's current stream and before its later work."""
        with torch.cuda.device(self.device):
            outer = torch.cuda.current_stream()
            if self.isolated or outer == self.stream:
                with torch.cuda.stream(self.stream):
                    yield
                return
            self.stream.wait_stream(outer)
            with torch.cuda.stream(self.stream):
                yield
            outer.wait_stream(self.stream)

    def synchronize(self):
        self._check(self.lib.datmo_synchronize(self.h))

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.tdev)

    def profile(self, on: bool, tags=None):
        """Bracket tagged launches with CUDA events; `tags` (names from _lib.TAGS) restricts it to those."""
        if on and tags is not None:
            mask = 0
            for t in tags:
                mask |= 1 << _lib.TAGS.index(t)
            self._check(self.lib.datmo_profile_tags(self.h, mask))
        else:
            self._check(self.lib.datmo_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._check(self.lib.datmo_profile_reset(self.h))

    def profile_read(self) -> dict:
        n = (C.c_int64 * _lib.TAG_COUNT)()
        ms = (C.c_double * _lib.TAG_COUNT)()
        self._check(self.lib.datmo_profile_read(self.h, n, ms))
        return {t: dict(launches=int(n[i]), ms=float(ms[i])) for i, t in enumerate(_lib.TAGS)}

    def launch_count(self) -> int:
        return int(self.lib.datmo_launch_count(self.h))

    def workspace_bytes(self) -> int:
        return int(self.lib.datmo_workspace_bytes(self.h))

    @staticmethod
    def _img_dtype(t: torch.Tensor) -> int:
        if t.dtype == torch.uint8:
            return _lib.U8
        if t.dtype == torch.float32:
            return _lib.F32
        raise TypeError("images must be uint8 or float32")

    def _batched(self, t: torch.Tensor):
        if t.dim() == 2:
            t = t.unsqueeze(0)
        if t.dim() != 3:
            raise ValueError("expected [H,W] or [B,H,W]")
        if t.device != self.tdev:
            raise ValueError(f"tensor is on {t.device}, engine on {self.tdev}")
        return t.contiguous()

    # -- Farneback ------------------------------------------------------------------------
    def farneback_layers(self, H: int, W: int, params=None):
        p = params or farneback_params()
        w = (C.c_int * 16)()
        h = (C.c_int * 16)()
        n = self.lib.datmo_farneback_layers(H, W, C.byref(p), 16, w, h)
        return [(h[i], w[i]) for i in range(n)]

    def farneback(self, prev: torch.Tensor, nxt: torch.Tensor, params=None, out: torch.Tensor | None = None):
        """cv2.calcOpticalFlowFarneback for a batch: [B,H,W] u8/f32 x2 -> [B,H,W,2] f32."""
        p = params or farneback_params()
        with self.on_stream():
            prev = self._batched(prev)
            nxt = self._batched(nxt)
            if prev.shape != nxt.shape or prev.dtype != nxt.dtype:
                raise ValueError("prev and next must have the same shape and dtype")
            B, H, W = prev.shape
            flow = out if out is not None else self.empty((B, H, W, 2), torch.float32)
            self._check(self.lib.datmo_farneback_dev(self.h, _ptr(prev), _ptr(nxt), self._img_dtype(prev), H, W, B,
                                                     C.byref(p), _ptr(flow)))
        return flow

    # stage-level entry points (parity tests diff these against the oracle)
    def fb_pyramid_image(self, img, ksize, sigma, h_out, w_out):
        with self.on_stream():
            img = self._batched(img)
            B, H, W = img.shape
            out = self.empty((B, h_out, w_out), torch.float32)
            self._check(self.lib.datmo_fb_pyramid_image_dev(self.h, _ptr(img), self._img_dtype(img), H, W, B, ksize,
                                                            float(sigma), h_out, w_out, _ptr(out)))
        return out

    def fb_polyexp(self, img, poly_n, poly_sigma):
        with self.on_stream():
            img = self._batched(img)
            B, H, W = img.shape
            R = self.empty((B, 5, H, W), torch.float32)
            self._check(self.lib.datmo_fb_polyexp_dev(self.h, _ptr(img), H, W, B, poly_n, float(poly_sigma), _ptr(R)))
        return R

    def fb_update_matrices(self, R0, R1, flow):
        with self.on_stream():
            B, _, H, W = R0.shape
            M = self.empty((B, 5, H, W), torch.float32)
            self._check(self.lib.datmo_fb_update_matrices_dev(self.h, _ptr(R0.contiguous()), _ptr(R1.contiguous()),
                                                              _ptr(flow.contiguous()), H, W, B, _ptr(M)))
        return M

    def fb_blur_solve(self, M, winsize):
        with self.on_stream():
            B, _, H, W = M.shape
            flow = self.empty((B, H, W, 2), torch.float32)
            self._check(self.lib.datmo_fb_blur_solve_dev(self.h, _ptr(M.contiguous()), H, W, B, winsize, _ptr(flow)))
        return flow

    def fb_flow_iter(self, R0, R1, flow, winsize):
        with self.on_stream():
            B, _, H, W = R0.shape
            out = self.empty((B, H, W, 2), torch.float32)
            self._check(self.lib.datmo_fb_flow_iter_dev(self.h, _ptr(R0.contiguous()), _ptr(R1.contiguous()),
                                                        _ptr(flow.contiguous()), H, W, B, winsize, _ptr(out)))
        return out

    def fb_upsample_flow(self, flow, h_out, w_out, mul):
        with self.on_stream():
            B, H, W, _ = flow.shape
            out = self.empty((B, h_out, w_out, 2), torch.float32)
            self._check(self.lib.datmo_fb_upsample_flow_dev(self.h, _ptr(flow.contiguous()), H, W, B, h_out, w_out,
                                                            float(mul), _ptr(out)))
        return out

    # -- velocity / masks ------------------------------------------------------------------
    def velocity_mask(self, flow: torch.Tensor, px_x: float, px_y: float, alpha_cont: float, thresh: float = 0.1,
                      want=("vx", "vy", "ang", "mask", "vx_f", "vy_f", "valid", "n_valid")) -> dict:
        """flow [B,H,W,2] -> dict of the requested grids (see datmo_velocity_mask_dev)."""
        with self.on_stream():
            if flow.dim() == 3:
                flow = flow.unsqueeze(0)
            flow = flow.contiguous()
            B, H, W, _ = flow.shape
            out = {}
            for k in ("vx", "vy", "ang", "vx_f", "vy_f", "ang_f"):
                out[k] = self.empty((B, H, W), torch.float32) if k in want else None
            for k in ("mask", "valid"):
                out[k] = self.empty((B, H, W), torch.uint8) if k in want else None
            out["n_valid"] = self.empty((B,), torch.int32) if "n_valid" in want else None
            if out["ang_f"] is not None:
                for k in ("vx_f", "vy_f"):
                    if out[k] is None:
                        out[k] = self.empty((B, H, W), torch.float32)
            self._check(self.lib.datmo_velocity_mask_dev(
                self.h, _ptr(flow), H, W, B, float(px_x), float(px_y), float(alpha_cont), float(thresh),
                _ptr(out["vx"]), _ptr(out["vy"]), _ptr(out["ang"]), _ptr(out["mask"]), _ptr(out["vx_f"]),
                _ptr(out["vy_f"]), _ptr(out["ang_f"]), _ptr(out["valid"]), _ptr(out["n_valid"])))
        return {k: v for k, v in out.items() if v is not None}

    def filtered_grids_f64(self, vx_f: torch.Tensor, vy_f: torch.Tensor, want=("vx", "vy", "mag", "ang")) -> dict:
        """float64 view of the filtered field, its magnitude and curl (main.py:600-606), one launch."""
        with self.on_stream():
            squeeze = vx_f.dim() == 2
            if squeeze:
                vx_f, vy_f = vx_f.unsqueeze(0), vy_f.unsqueeze(0)
            vx_f, vy_f = vx_f.contiguous(), vy_f.contiguous()
            B, H, W = vx_f.shape
            out = {k: (self.empty((B, H, W), torch.float64) if k in want else None) for k in ("vx", "vy", "mag", "ang")}
            self._check(self.lib.datmo_filtered_grids_f64_dev(self.h, _ptr(vx_f), _ptr(vy_f), H, W, B, _ptr(out["vx"]),
                                                              _ptr(out["vy"]), _ptr(out["mag"]), _ptr(out["ang"])))
        return {k: (v[0] if squeeze else v) for k, v in out.items() if v is not None}

    def narrow_f64(self, t: torch.Tensor) -> torch.Tensor:
        """float64 -> float32 on the device; raises when a value is not float32-representable."""
        with self.on_stream():
            t = t.contiguous()
            out = self.empty(t.shape, torch.float32)
            lossy = C.c_int(0)
            self._check(self.lib.datmo_narrow_f64_dev(self.h, _ptr(t), t.numel(), _ptr(out), C.byref(lossy)))
        if lossy.value:
            raise ValueError("velocities must be float32-representable (they are f32 flow * mask in the reference)")
        return out

    def propagation_mask(self, vx: torch.Tensor, vy: torch.Tensor, dt: float, grid_resolution, alpha_p: float,
                         ax: torch.Tensor | None = None, ay: torch.Tensor | None = None) -> torch.Tensor:
        """[B,H,W] (or [H,W]) f32 / f64 velocities (and accelerations) -> uint8 mask, see datmo_propagation_mask_dev."""
        with self.on_stream():
            squeeze = vx.dim() == 2
            ts = [t if t is None else (t.unsqueeze(0) if squeeze else t) for t in (vx, vy, ax, ay)]
            dtype = torch.float64 if vx.dtype == torch.float64 else torch.float32
            ts = [t if t is None else t.to(dtype).contiguous() for t in ts]
            B, H, W = ts[0].shape
            mask = self.empty((B, H, W), torch.uint8)
            self._check(self.lib.datmo_propagation_mask_dev(
                self.h, _ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]), _ptr(ts[3]), 2 if dtype == torch.float64 else 1, H, W, B,
                float(dt), float(grid_resolution[0]), float(grid_resolution[1]), float(alpha_p), _ptr(mask)))
        return mask[0] if squeeze else mask

    # -- DBSCAN ---------------------------------------------------------------------------------
    def dbscan_grid(self, vx_f, vy_f, valid, eps: float, min_samples: int, cap: int | None = None):
        """-> (n_valid [B] i32, labels [B,cap] i32, indices [B,cap,2] i32, n_clusters [B] i32), device."""
        with self.on_stream():
            if vx_f.dim() == 2:
                vx_f, vy_f, valid = vx_f.unsqueeze(0), vy_f.unsqueeze(0), valid.unsqueeze(0)
            vx_f, vy_f = vx_f.contiguous(), vy_f.contiguous()
            valid = valid.to(torch.uint8).contiguous()
            B, H, W = vx_f.shape
            cap = H * W if cap is None else int(cap)
            n_valid = self.empty((B,), torch.int32)
            n_clusters = self.empty((B,), torch.int32)
            labels = self.empty((B, cap), torch.int32)
            indices = self.empty((B, cap, 2), torch.int32)
            self._check(self.lib.datmo_dbscan_grid_dev(self.h, _ptr(vx_f), _ptr(vy_f), _ptr(valid), H, W, B,
                                                       float(eps), int(min_samples), cap, _ptr(n_valid), _ptr(labels),
                                                       _ptr(indices), _ptr(n_clusters)))
        return n_valid, labels, indices, n_clusters

    def cluster_summary(self, vx_f, vy_f, n_valid, labels, indices, max_clusters: int):
        """-> [B,max_clusters,8] f64: n, mean row, mean col, mean vx, mean vy, cov rr, rc, cc."""
        with self.on_stream():
            if vx_f.dim() == 2:
                vx_f, vy_f = vx_f.unsqueeze(0), vy_f.unsqueeze(0)
            B, H, W = vx_f.shape
            cap = labels.shape[1]
            out = self.empty((B, max_clusters, 8), torch.float64)
            self._check(self.lib.datmo_cluster_summary_dev(self.h, _ptr(vx_f.contiguous()), _ptr(vy_f.contiguous()), H,
                                                           W, B, cap, _ptr(n_valid), _ptr(labels), _ptr(indices),
                                                           int(max_clusters), _ptr(out)))
        return out

    def pack_indices(self, indices: torch.Tensor, n_valid: torch.Tensor) -> torch.Tensor:
        """[B,cap,2] int32 (row, col) -> [B,cap] int32 holding (row << 16) | col for the first n_valid cells
        of every frame (the rest is left untouched): a third less to read back than the pairs."""
        with self.on_stream():
            B, cap, _ = indices.shape
            packed = self.empty((B, cap), torch.int32)
            self._check(self.lib.datmo_pack_indices_dev(self.h, _ptr(indices.contiguous()), _ptr(n_valid), cap, B,
                                                        _ptr(packed)))
        return packed

    # -- flow -> clusters, the body of the reference's driver loop (main.py:577-615) --------------
    def flow_pipeline(self, prev, nxt, px_x, px_y, alpha_cont, eps, min_samples, params=None, thresh=0.1,
                      cap: int | None = None, max_clusters: int = 0, keep_flow: bool = True,
                      flow_buf: torch.Tensor | None = None, want_ang_f: bool = False) -> FlowPipelineResult:
        with self.on_stream():
            flow = self.farneback(prev, nxt, params, out=flow_buf)
            want = ("vx_f", "vy_f", "valid", "ang_f") if want_ang_f else ("vx_f", "vy_f", "valid")
            vm = self.velocity_mask(flow, px_x, px_y, alpha_cont, thresh, want=want)
            n_valid, labels, indices, n_clusters = self.dbscan_grid(vm["vx_f"], vm["vy_f"], vm["valid"], eps,
                                                                    min_samples, cap)
            summary = None
            if max_clusters > 0:
                summary = self.cluster_summary(vm["vx_f"], vm["vy_f"], n_valid, labels, indices, max_clusters)
        return FlowPipelineResult(flow if keep_flow else None, vm["vx_f"], vm["vy_f"], vm["valid"], n_valid, labels,
                                  indices, n_clusters, summary, labels.shape[1], vm.get("ang_f"))

    # -- BEV / RANSAC / preprocessing -----------------------------------------------------------
    def bev_bins(self, lo: float, hi: float, step: float) -> int:
        return int(self.lib.datmo_bev_bins(float(lo), float(hi), float(step)))

    @staticmethod
    def _points_layout(points: torch.Tensor) -> int:
        if points.dtype == torch.float64 and points.dim() == 2 and points.shape[1] == 3:
            return _lib.PTS_F64_XYZ
        if points.dtype == torch.float32 and points.dim() == 2 and points.shape[1] == 4:
            return _lib.PTS_F32_XYZW
        raise TypeError("points must be float64 [n,3] or float32 [n,4]")

    def bev_rasterize(self, points: torch.Tensor, grid_resolution, x_range, y_range, a=0.5, b=0.5, h_max=5.0):
        """compute_bev_grid on the device: points f64 [n,3] or f32 [n,4] -> uint8 [nx,ny]."""
        with self.on_stream():
            points = points.contiguous()
            layout = self._points_layout(points)
            nx = self.bev_bins(x_range[0], x_range[1], grid_resolution[0])
            ny = self.bev_bins(y_range[0], y_range[1], grid_resolution[1])
            bev = self.empty((nx, ny), torch.uint8)
            self._check(self.lib.datmo_bev_rasterize_dev(self.h, _ptr(points), layout, points.shape[0],
                                                         float(grid_resolution[0]), float(grid_resolution[1]),
                                                         float(x_range[0]), float(y_range[0]), nx, ny, float(a),
                                                         float(b), float(h_max), _ptr(bev)))
        return bev

    def roi_filter(self, points: torch.Tensor, roi_bounds):
        """filter_points_in_roi on the device (stable compaction) -> points inside the closed ROI box."""
        with self.on_stream():
            points = points.contiguous()
            layout = self._points_layout(points)
            out = torch.empty_like(points)
            roi = (C.c_double * 6)(*[float(v) for v in roi_bounds])
            n_out = C.c_int64(0)
            self._check(self.lib.datmo_roi_filter_dev(self.h, _ptr(points), layout, points.shape[0], roi, _ptr(out),
                                                      C.byref(n_out)))
        return out[:n_out.value]

    def expand_points(self, points: torch.Tensor, expansion: int, noise_std: float = 0.01,
                      noise: torch.Tensor | None = None, seed: int = 0):
        """increase_point_density on the device: f64 [n,3] -> f64 [n*expansion,3]."""
        with self.on_stream():
            points = points.to(torch.float64).contiguous()
            n = points.shape[0]
            out = self.empty((n * expansion, 3), torch.float64)
            if noise is not None:
                noise = noise.to(torch.float64).contiguous()
                if noise.numel() != n * expansion * 3:
                    raise ValueError("noise must have the shape of the expanded cloud")
            self._check(self.lib.datmo_expand_points_dev(self.h, _ptr(points), n, int(expansion), float(noise_std),
                                                         _ptr(noise), C.c_uint64(seed), _ptr(out)))
        return out

    def ransac_ground(self, points: torch.Tensor, distance_threshold=0.5, ransac_n=5, num_iterations=5000, seed=0,
                      flip_x=False, return_hypotheses=False):
        """segment_plane on the device -> dict(plane, refit, inlier_mask, best[, hyp_*])."""
        with self.on_stream():
            points = points.contiguous()
            layout = self._points_layout(points)
            n = points.shape[0]
            plane = self.empty((4,), torch.float64)
            refit = self.empty((4,), torch.float64)
            mask = self.empty((n,), torch.uint8)
            best = self.empty((2,), torch.int32)
            hp = hc = he = None
            if return_hypotheses:
                hp = self.empty((num_iterations, 4), torch.float64)
                hc = self.empty((num_iterations,), torch.int32)
                he = self.empty((num_iterations,), torch.float64)
            self._check(self.lib.datmo_ransac_ground_dev(self.h, _ptr(points), layout, n, int(flip_x),
                                                         float(distance_threshold), int(ransac_n),
                                                         int(num_iterations), C.c_uint64(seed), _ptr(plane),
                                                         _ptr(refit), _ptr(mask), _ptr(best), _ptr(hp), _ptr(hc),
                                                         _ptr(he)))
        out = dict(plane=plane, refit=refit, inlier_mask=mask, best=best)
        if return_hypotheses:
            out.update(hyp_planes=hp, hyp_count=hc, hyp_err=he)
        return out

    def preprocess(self, points_xyzw: torch.Tensor, grid_resolution, x_range, y_range, z_max, roi_bounds,
                   distance_threshold=0.5, ransac_n=5, num_iterations=5000, seed=0, flip_x=True,
                   ground_mask: torch.Tensor | None = None, expansion=10, noise_std=0.01,
                   noise: torch.Tensor | None = None):
        """preprocess_pcd after the file read (main.py:65-92) -> uint8 [nx,ny] or None when the ROI is empty."""
        with self.on_stream():
            pts = points_xyzw.contiguous()
            if self._points_layout(pts) != _lib.PTS_F32_XYZW:
                raise TypeError("preprocess takes float32 [n,4] points")
            nx = self.bev_bins(x_range[0], x_range[1], grid_resolution[0])
            ny = self.bev_bins(y_range[0], y_range[1], grid_resolution[1])
            bev = self.empty((nx, ny), torch.uint8)
            roi = (C.c_double * 6)(*[float(v) for v in roi_bounds])
            n_roi = C.c_int64(0)
            if noise is not None:
                noise = noise.to(torch.float64).contiguous()
                if noise.numel() != pts.shape[0] * expansion * 3:
                    raise ValueError("noise must be [n, expansion, 3]")
            if ground_mask is not None:
                ground_mask = ground_mask.to(torch.uint8).contiguous()
            st = self.lib.datmo_preprocess_dev(self.h, _ptr(pts), pts.shape[0], int(flip_x), float(distance_threshold),
                                               int(ransac_n), int(num_iterations), C.c_uint64(seed), _ptr(ground_mask),
                                               roi, int(expansion), float(noise_std), _ptr(noise),
                                               float(grid_resolution[0]), float(grid_resolution[1]), float(x_range[0]),
                                               float(y_range[0]), nx, ny, float(z_max), _ptr(bev), C.byref(n_roi))
            if st == _lib.E_EMPTY:
                return None
            self._check(st)
        return bev


_default: dict[int, Engine] = {}


def default_engine(device: int | None = None) -> Engine:
    if not torch.cuda.is_available():
        raise _lib.DatmoLibraryError("no CUDA device: datmo_b200 has no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    if dev not in _default:
        _default[dev] = Engine(dev)
    return _default[dev]


def chain_config(H: int, W: int, batch: int, px_x: float, px_y: float, alpha_cont: float, eps: float,
                 min_samples: int, cap: int, max_clusters: int = 1024, params=None, dtype=_lib.U8, thresh: float = 0.1,
                 want_cells: bool = True, n_slots: int = 2) -> _lib.ChainConfig:
    """datmo_chain_config with the reference's defaults filled in by the library."""
    cfg = _lib.ChainConfig()
    _lib.load().datmo_chain_default_config(C.byref(cfg))
    cfg.H, cfg.W, cfg.batch, cfg.dtype = H, W, batch, dtype
    cfg.px_x, cfg.px_y, cfg.alpha_cont, cfg.thresh = float(px_x), float(px_y), float(alpha_cont), float(thresh)
    cfg.eps, cfg.min_samples = float(eps), int(min_samples)
    cfg.cap, cfg.max_clusters, cfg.want_cells, cfg.n_slots = int(cap), int(max_clusters), int(want_cells), int(n_slots)
    if params is not None:
        cfg.fb = params
    return cfg


class HostFlowPipeline:
    """Host-buffer entry to the flow -> clusters chain for a stream of frame-pair batches: a thin
    caller of the C-ABI chain object (datmo_chain_create / _submit / _collect, include/datmo_b200.h).

    Host uint8 (or float32) pairs go in, host arrays come out (counts, labels, cell indices, cluster
    summaries — what the reference's driver loop consumes after main.py:577-615).  The library
    double-buffers the copies against the kernels on its own streams and reads every array back with
    ONE device-to-host copy per batch.  Usage: ``submit(slot, prev, nxt)`` then ``collect(slot)``;
    keep at most ``n_slots`` submissions outstanding.  Frames in pinned memory make the uploads
    asynchronous.
    """

    def __init__(self, eng: Engine, batch: int, H: int, W: int, px_x: float, px_y: float, alpha_cont: float,
                 eps: float, min_samples: int, params=None, cap: int | None = None, max_clusters: int = 1024,
                 n_slots: int = 2, want_cells: bool = True, dtype=_lib.U8):
        self.eng, self.B, self.H, self.W = eng, batch, H, W
        self.cap = H * W if cap is None else int(cap)
        self.max_clusters = max_clusters
        self.cfg = chain_config(H, W, batch, px_x, px_y, alpha_cont, eps, min_samples, self.cap, max_clusters, params,
                                dtype=dtype, want_cells=want_cells, n_slots=n_slots)
        c = C.c_void_p()
        eng._check(eng.lib.datmo_chain_create(eng.h, C.byref(self.cfg), C.byref(c)))
        self.c = c
        self._held = [None] * n_slots     # the caller's frames stay alive until the slot is collected
        self.h2d_bytes = 2 * batch * H * W * (1 if dtype == _lib.U8 else 4)
        self.d2h_bytes = 0

    def close(self):
        if getattr(self, "c", None) and getattr(self.eng, "h", None):
            self.eng.lib.datmo_chain_destroy(self.c)
        self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st != _lib.OK:
            raise _lib.DatmoError(st, self.eng.lib.datmo_chain_last_error(self.c).decode(errors="replace"))

    @staticmethod
    def _host_ptr(a):
        if isinstance(a, torch.Tensor):
            if a.is_cuda or not a.is_contiguous():
                raise ValueError("frames must be contiguous host tensors")
            return a.data_ptr()
        a = np.ascontiguousarray(a)
        return a.ctypes.data

    def submit(self, slot: int, prev_host, next_host):
        """prev_host / next_host: [B,H,W] host arrays (numpy or torch, ideally pinned)."""
        self._held[slot] = (prev_host, next_host)
        self._check(self.eng.lib.datmo_chain_submit(self.c, slot, C.c_void_p(self._host_ptr(prev_host)),
                                                    C.c_void_p(self._host_ptr(next_host))))

    def collect(self, slot: int):
        """-> (n_valid[B], n_clusters[B], offsets[B+1], labels[sum n], cells[sum n], summary[B,kmax,8]):
        numpy views into the chain's pinned buffers (valid until the slot is submitted again).  The cells of
        pair b are rows offsets[b]:offsets[b+1].  labels is int16 when every pair of the batch has fewer than
        32768 clusters, else int32; cells is uint32 (row << 16) | col (``unpack_indices`` gives [sum n, 2])."""
        r = _lib.ChainResult()
        self._check(self.eng.lib.datmo_chain_collect(self.c, slot, C.byref(r)))
        self._held[slot] = None
        B = self.B
        n_valid = np.ctypeslib.as_array(r.n_valid, shape=(B,))
        n_clusters = np.ctypeslib.as_array(r.n_clusters, shape=(B,))
        offsets = np.ctypeslib.as_array(r.offsets, shape=(B + 1,))
        total = int(offsets[B])
        if r.labels and total:
            lt = C.c_int16 if r.label_bytes == 2 else C.c_int32
            labels = np.ctypeslib.as_array(C.cast(r.labels, C.POINTER(lt)), shape=(total,))
            cells = np.ctypeslib.as_array(r.cells, shape=(total,))
        else:
            labels, cells = np.zeros(0, np.int16), np.zeros(0, np.uint32)
        k = r.summary_rows
        summary = np.ctypeslib.as_array(r.summary, shape=(B, k, 8)) if k else np.zeros((B, 0, 8))
        self.d2h_bytes = int(r.d2h_bytes)
        self.truncated = bool(r.truncated)
        return n_valid, n_clusters, offsets, labels, cells, summary

    @staticmethod
    def unpack_indices(packed: np.ndarray) -> np.ndarray:
        """[n] (row << 16) | col -> [n, 2] int32 (row, col), as dbscan_clustering's valid_indices."""
        u = packed.view(np.uint32)
        return np.stack([(u >> 16).astype(np.int32), (u & 0xffff).astype(np.int32)], axis=1)
