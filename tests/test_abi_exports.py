"""CPU: the C-ABI library builds, loads and exports every symbol include/datmo_b200.h declares.
No compute call is made (there is no GPU here)."""
import os
import re

import pytest

from datmo_using_optical_flow_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "datmo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(datmo_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert lib.datmo_abi_version() == _lib.ABI_VERSION == 2


def test_host_side_helpers_without_gpu():
    lib = _lib.load()
    p = _lib.FarnebackParams()
    lib.datmo_farneback_default_params(p)
    assert (p.pyr_scale, p.levels, p.winsize, p.iterations, p.poly_n, p.poly_sigma, p.flags) == (0.3, 5, 15, 5, 5, 5.0, 0)
    import ctypes as C
    w = (C.c_int * 8)()
    h = (C.c_int * 8)()
    n = lib.datmo_farneback_layers(1024, 1024, p, 8, w, h)
    assert n == 3 and list(w[:3]) == [92, 307, 1024] and list(h[:3]) == [92, 307, 1024]
    import numpy as np
    for lo, hi, st in [(-50, 50, 0.25), (-50, 50, 0.125), (-51.2, 51.2, 0.1), (-51.2, 51.2, 0.05), (-20, 20, 0.2), (0, 1, 0.3)]:
        assert lib.datmo_bev_bins(lo, hi, st) == len(np.arange(lo, hi, st))


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from datmo_using_optical_flow_b200 import main
    import numpy as np
    with pytest.raises(_lib.DatmoLibraryError):
        main.compute_bev_grid(np.zeros((4, 3)), (0.5, 0.5), (-1, 1), (-1, 1))
    import ctypes as C
    h = C.c_void_p()
    assert _lib.load().datmo_create(0, None, C.byref(h)) != 0


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "datmo_using_optical_flow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
