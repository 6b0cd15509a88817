"""Import the UNMODIFIED reference ``main.py`` in the build container — TEST ORACLE.

/root/reference/Optical_flow/main.py imports open3d, matplotlib and shapely at
module level (main.py:2-10); none is installed in this image.  Stub modules are
injected so the import succeeds; every function that does not touch them
(compute_bev_grid, compute_velocity_vectors, continuity_mask,
dbscan_clustering, extract_cluster_data, EKF, track_clusters, ...) then runs
unmodified against the live cv2 / sklearn / numpy.  ``preprocess_pcd`` cannot
run (needs real Open3D I/O + RANSAC).

Importing main.py creates an output directory in the CWD (main.py:21-23), so
the import happens from a scratch directory.  This module is used only by
tests/golden/make_golden.py and by CPU tests that skip when /root/reference is
absent (it does not exist on the GPU box).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
import types

REFERENCE_DIR = "/root/reference/Optical_flow"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "main.py"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


_cached = None


def load_reference_main():
    """Returns the reference's ``main`` module (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(REFERENCE_DIR)
    o3d = _stub("open3d")
    for sub in ("io", "geometry", "utility"):
        setattr(o3d, sub, _stub(f"open3d.{sub}"))
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.cm = _stub("matplotlib.cm")
    mpl.path = _stub("matplotlib.path", Path=object)
    shp = _stub("shapely")
    shp.geometry = _stub("shapely.geometry", Polygon=object, Point=object)
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="datmo_ref_")
    sys.path.insert(0, REFERENCE_DIR)
    try:
        os.chdir(scratch)
        import main as ref_main  # noqa: E402  (the reference's module)
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE_DIR)
    _cached = ref_main
    return ref_main


@contextlib.contextmanager
def quiet():
    """The reference prints whole arrays (main.py:141-161); silence it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
