"""Builds libdatmo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m datmo_using_optical_flow_b200.build [--force]

The shared library travels to the GPU box with the repo snapshot; it is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdatmo_b200.so")
STAMP = os.path.join(LIB_DIR, "libdatmo_b200.stamp")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["context.cu", "farneback.cu", "velmask.cu", "dbscan.cu", "dbscan_runs.cu", "bev.cu", "ransac.cu", "chain.cu"]
# files whose fp64 arithmetic must round every operation like numpy does (no FMA contraction)
NO_FMAD = {"ransac.cu"}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", INCLUDE]


class NvccMissing(RuntimeError):
    """The sources changed (or the library is absent) and there is no nvcc to build it."""


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise NvccMissing("nvcc not found: libdatmo_b200 cannot be built (there is no CPU fallback)")


def stamp_matches() -> bool:
    """True when the built library's stamp equals the digest of the sources in this tree."""
    try:
        with open(STAMP) as fh:
            return os.path.exists(LIB_PATH) and fh.read().strip() == _digest()
    except OSError:
        return False


def _digest() -> str:
    hsh = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    files += [os.path.join(INCLUDE, "datmo_b200.h"), __file__]
    for path in files:
        with open(path, "rb") as fh:
            hsh.update(os.path.basename(path).encode())   # path-independent: the prebuilt .so travels
            hsh.update(fh.read())
    return hsh.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources changed; returns the path of the shared library."""
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    # several ranks may import at once: one builds, the others wait and re-check
    import fcntl
    lock = open(os.path.join(LIB_DIR, ".build.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
            with open(STAMP) as fh:
                if fh.read().strip() == digest:
                    return LIB_PATH
        return _build_locked(nvcc, digest, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(nvcc: str, digest: str, verbose: bool) -> str:
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, "-c", os.path.join(CSRC, src), "-o", obj]
        if src in NO_FMAD:
            cmd.insert(1, "-fmad=false")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, cmd, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{' '.join(cmd)}\n{out}\n")
        elif verbose and out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("building libdatmo_b200.so failed")
    link = [nvcc, *ARCH, "-shared", "-o", LIB_PATH, *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"linking libdatmo_b200.so failed:\n{res.stdout}")
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
