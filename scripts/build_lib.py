"""Build libminidiff_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the
library is a plain C-ABI shared object, see include/minidiff_b200.h).

    python scripts/build_lib.py            # incremental
    python scripts/build_lib.py --force

Kept OUTSIDE the package on purpose: importing `minidiff_b200` loads the shared library, so the
thing that creates the library must not need that import.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "minidiff_b200")
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "lib", "libminidiff_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
    "-diag-suppress", "550",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; minidiff_b200 needs the CUDA toolkit to build")


def _digest(paths) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_library(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "minidiff_b200.h"))
    nvcc = _nvcc()
    jobs = []
    for src in sources():
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest([path] + headers)
        fresh = (not force and os.path.exists(obj) and os.path.exists(stamp)
                 and open(stamp).read() == dig)
        if not fresh:
            jobs.append((path, obj, stamp, dig))

    def compile_one(job):
        path, obj, stamp, dig = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {path}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as fh:
            fh.write(dig)
        return path

    if jobs:
        if verbose:
            print(f"[build_lib] compiling {len(jobs)} source(s) for sm_100a ...", flush=True)
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for done in ex.map(compile_one, jobs):
                if verbose:
                    print("  built", os.path.relpath(done, ROOT), flush=True)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources()]
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                    "-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print("[build_lib] linked", os.path.relpath(LIB, ROOT), flush=True)
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv)
