"""Build libgstex_b200.so in-tree with nvcc for sm_100a (no torch headers; seconds, not minutes).

    python -m gstex_cuda_b200.build [--force] [--verbose]

Each .cu is compiled to an object in csrc/_build/ (in parallel) and linked into
gstex_cuda_b200/libgstex_b200.so.  `-Xptxas -v` output is kept in csrc/_build/ptxas.log.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libgstex_b200.so")
SOURCES = ["util.cu", "project.cu", "binning.cu", "binning_tiles.cu", "pack.cu", "raster_forward.cu", "raster_backward.cu", "sh.cu",
           "texture_sample.cu", "texture_edit.cu", "train_ops.cu", "loss.cu", "pipeline.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v", "--expt-relaxed-constexpr"]


# experiments: extra flags (e.g. -DGSTEX_EXP_...) through the environment, never set in production
NVCC_FLAGS += os.environ.get("GSTEX_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "gstex_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile and link in-tree.  Safe to call from several processes at once (all ranks of a torchrun launch import the
    package together): an exclusive file lock serialises them, the library is linked under a temporary name and renamed
    into place, and whoever gets the lock second finds the library up to date."""
    import fcntl

    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    nvcc = _nvcc()
    hdr_mtime = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    hdr_mtime = max(hdr_mtime, os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "gstex_b200.h")))

    def compile_one(src: str):
        obj = os.path.join(BUILD, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), hdr_mtime):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, f"==== {src}\n{r.stderr}"

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if log:
        with open(os.path.join(BUILD, "ptxas.log"), "w") as f:  # the log of the LAST build (git-ignored)
            f.write(log)
        if verbose:
            print(log)
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)  # atomic: no process ever dlopens a half-written library
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
