"""In-tree build of libisac_b200.so (hand-written CUDA for sm_100a, C ABI in include/isac_b200.h).

nvcc cross-compiles without a GPU; the .so stays in-tree (git-ignored, shipped by gpurun).
Usage: ``python -m <package>.build`` or ``build_library()``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libisac_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xptxas", "-v",
    "--fmad=true",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libisac_b200.so cannot be built")
    return nvcc


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest() -> str:
    h = hashlib.sha1()
    for root in (CSRC, INCLUDE_DIR):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc, src, obj, log):
    cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE_DIR, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return src


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link lib/libisac_b200.so. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = _headers_digest()
    stamp = os.path.join(OBJ_DIR, "headers.sha1")
    old = open(stamp).read() if os.path.exists(stamp) else ""
    headers_changed = old != digest
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(obj)
        stale = (force or headers_changed or not os.path.exists(obj)
                 or os.path.getmtime(obj) < os.path.getmtime(os.path.join(CSRC, src)))
        if stale:
            jobs.append((src, obj, os.path.join(OBJ_DIR, src[:-3] + ".log")))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for done in ex.map(lambda j: _compile_one(nvcc, *j), jobs):
                if verbose:
                    print(f"[isac_b200.build] compiled {done}", file=sys.stderr)
        with open(stamp, "w") as fh:
            fh.write(digest)
    need_link = bool(jobs) or not os.path.exists(LIB_PATH)
    if need_link:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs,
               "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[isac_b200.build] linked {LIB_PATH}", file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
