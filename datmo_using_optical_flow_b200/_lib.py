"""ctypes binding of libdatmo_b200.so (the C ABI declared in include/datmo_b200.h).

There is no CPU fallback: if the library cannot be loaded (and cannot be built
because nvcc is absent) importing a compute function raises ``DatmoLibraryError``.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ABI_VERSION = 2   # DATMO_ABI_VERSION of include/datmo_b200.h these bindings were written against

OK = 0
E_INVALID = -1
E_CUDA = -2
E_CAPACITY = -3
E_EMPTY = -4

U8 = 0
F32 = 1
PTS_F64_XYZ = 0
PTS_F32_XYZW = 1

TAGS = ["pyramid", "polyexp", "flow_init", "flow_iter", "velmask", "dbscan", "bev", "ransac", "cluster"]
TAG_COUNT = len(TAGS)


class DatmoLibraryError(RuntimeError):
    pass


class DatmoError(RuntimeError):
    """A C-ABI call returned a negative status (mirrors the reference's habit of
    raising plain exceptions that process_multiple_frames catches, main.py:635-637)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"datmo_b200 status {status}: {message}")
        self.status = status


class FarnebackParams(C.Structure):
    _fields_ = [("pyr_scale", C.c_double), ("levels", C.c_int), ("winsize", C.c_int), ("iterations", C.c_int),
                ("poly_n", C.c_int), ("poly_sigma", C.c_double), ("flags", C.c_int), ("variant", C.c_int)]


class ChainConfig(C.Structure):
    """datmo_chain_config (include/datmo_b200.h)."""
    _fields_ = [("H", C.c_int), ("W", C.c_int), ("batch", C.c_int), ("dtype", C.c_int), ("px_x", C.c_double),
                ("px_y", C.c_double), ("alpha_cont", C.c_double), ("thresh", C.c_double), ("eps", C.c_double),
                ("min_samples", C.c_int), ("cap", C.c_int), ("max_clusters", C.c_int), ("want_cells", C.c_int),
                ("n_slots", C.c_int), ("fb", FarnebackParams)]


class ChainResult(C.Structure):
    """datmo_chain_result (include/datmo_b200.h)."""
    _fields_ = [("n_valid", C.POINTER(C.c_int32)), ("n_clusters", C.POINTER(C.c_int32)),
                ("offsets", C.POINTER(C.c_int64)), ("labels", C.c_void_p), ("label_bytes", C.c_int),
                ("cells", C.POINTER(C.c_uint32)), ("summary", C.POINTER(C.c_double)), ("summary_rows", C.c_int),
                ("truncated", C.c_int), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


_vp, _i, _i64, _d, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_uint64
_pp = C.POINTER(FarnebackParams)
_cp = C.POINTER(ChainConfig)

# name -> (restype, argtypes); every symbol include/datmo_b200.h declares
SIGNATURES = {
    "datmo_create": (_i, [_i, _vp, C.POINTER(_vp)]),
    "datmo_destroy": (_i, [_vp]),
    "datmo_abi_version": (_i, []),
    "datmo_last_error": (C.c_char_p, [_vp]),
    "datmo_synchronize": (_i, [_vp]),
    "datmo_workspace_bytes": (C.c_size_t, [_vp]),
    "datmo_profile_enable": (_i, [_vp, _i]),
    "datmo_profile_tags": (_i, [_vp, C.c_uint]),
    "datmo_profile_reset": (_i, [_vp]),
    "datmo_profile_read": (_i, [_vp, C.POINTER(_i64), C.POINTER(_d)]),
    "datmo_launch_count": (_i64, [_vp]),
    "datmo_farneback_default_params": (None, [_pp]),
    "datmo_farneback_layers": (_i, [_i, _i, _pp, _i, C.POINTER(_i), C.POINTER(_i)]),
    "datmo_farneback_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _pp, _vp]),
    "datmo_farneback_host": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _pp, _vp]),
    "datmo_fb_pyramid_image_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _d, _i, _i, _vp]),
    "datmo_fb_polyexp_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp]),
    "datmo_fb_update_matrices_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "datmo_fb_blur_solve_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "datmo_fb_flow_iter_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "datmo_fb_upsample_flow_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _d, _vp]),
    "datmo_velocity_mask_dev": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _d, _d] + [_vp] * 9),
    "datmo_filtered_grids_f64_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "datmo_narrow_f64_dev": (_i, [_vp, _vp, _i64, _vp, C.POINTER(_i)]),
    "datmo_propagation_mask_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _d, _d, _vp]),
    "datmo_dbscan_grid_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _i, _i, _vp, _vp, _vp, _vp]),
    "datmo_pack_indices_dev": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "datmo_cluster_summary_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "datmo_bev_bins": (_i, [_d, _d, _d]),
    "datmo_bev_rasterize_dev": (_i, [_vp, _vp, _i, _i64, _d, _d, _d, _d, _i, _i, _d, _d, _d, _vp]),
    "datmo_bev_rasterize_host": (_i, [_vp, _vp, _i, _i64, _d, _d, _d, _d, _i, _i, _d, _d, _d, _vp]),
    "datmo_roi_filter_dev": (_i, [_vp, _vp, _i, _i64, C.POINTER(_d), _vp, C.POINTER(_i64)]),
    "datmo_expand_points_dev": (_i, [_vp, _vp, _i64, _i, _d, _vp, _u64, _vp]),
    "datmo_ransac_ground_dev": (_i, [_vp, _vp, _i, _i64, _i, _d, _i, _i, _u64] + [_vp] * 7),
    "datmo_preprocess_dev": (_i, [_vp, _vp, _i64, _i, _d, _i, _i, _u64, _vp, C.POINTER(_d), _i, _d, _vp, _d, _d, _d,
                                  _d, _i, _i, _d, _vp, C.POINTER(_i64)]),
    "datmo_chain_default_config": (None, [_cp]),
    "datmo_chain_create": (_i, [_vp, _cp, C.POINTER(_vp)]),
    "datmo_chain_destroy": (_i, [_vp]),
    "datmo_chain_last_error": (C.c_char_p, [_vp]),
    "datmo_chain_submit": (_i, [_vp, _i, _vp, _vp]),
    "datmo_chain_collect": (_i, [_vp, _i, C.POINTER(ChainResult)]),
    "datmo_flow_to_clusters_host": (_i, [_vp, _vp, _vp, _cp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
}

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first when the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if build_if_missing:
        try:
            path = _build.build()
        except _build.NvccMissing as exc:  # no nvcc on this machine: the prebuilt file that travelled with the repo
            if not os.path.exists(path):
                raise DatmoLibraryError(
                    f"libdatmo_b200.so is missing and could not be built ({exc}); "
                    "run `python -m datmo_using_optical_flow_b200.build` on a machine with nvcc. "
                    "There is no CPU fallback.") from exc
            if not _build.stamp_matches():
                import warnings
                warnings.warn("libdatmo_b200.so was built from different sources than the ones in this tree "
                              "and nvcc is not available to rebuild it", RuntimeWarning)
        # a compile or link failure of edited sources propagates: running a stale library against newer
        # bindings is worse than failing
    if not os.path.exists(path):
        raise DatmoLibraryError(f"{path} not found; there is no CPU fallback")
    try:
        lib = C.CDLL(path)
    except OSError as exc:
        raise DatmoLibraryError(f"cannot load {path}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise DatmoLibraryError(f"{path} does not export {name}") from exc
        fn.restype = res
        fn.argtypes = args
    if lib.datmo_abi_version() != ABI_VERSION:
        raise DatmoLibraryError(f"{path} has ABI version {lib.datmo_abi_version()}, these bindings need {ABI_VERSION}: "
                                "rebuild with `python -m datmo_using_optical_flow_b200.build --force`")
    _lib = lib
    return lib
