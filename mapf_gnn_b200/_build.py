"""Compile the CUDA sources under ``csrc/`` into ``lib/libmapf_b200.so`` (sm_100a only).

``nvcc`` cross-compiles without a GPU, so this runs in the build container; the resulting ``.so``
is git-ignored but travels to the GPU box with the repo snapshot.  There is no CPU fallback: if the
library is missing and cannot be built, importing the ops fails.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmapf_b200.so")
STAMP = os.path.join(LIBDIR, "libmapf_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]
# development only (e.g. "-DMAPF_ENC_DEBUG" for the phase-timing hooks used by tools/enc_dbg.py); part of the digest
NVCC_FLAGS += os.environ.get("MAPF_B200_NVCC_EXTRA", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(ROOT, "include", "mapf_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())      # not the absolute path: the repo lives elsewhere on the GPU box
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(a for a in NVCC_FLAGS if not os.path.isabs(a)).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_fresh():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Build (if stale) and return the path of the shared library.  Safe under concurrent ranks: an exclusive file
    lock serialises builders and the library is moved into place atomically."""
    override = os.environ.get("MAPF_B200_LIB_OVERRIDE")     # development: an alternative build (tools/enc_tc_variants.py)
    if override:
        return override
    if not force and is_fresh():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        if os.path.exists(LIB):       # GPU box without a toolchain on PATH: use what travelled
            return LIB
        raise RuntimeError("nvcc not found and libmapf_b200.so is not built")
    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_fresh():          # another rank built it while we waited
                return LIB
            return _build_locked(nvcc, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(nvcc, verbose):
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if pr.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        sys.stderr.write(out.stdout)
        raise RuntimeError("link failed")
    os.replace(tmp, LIB)                          # atomic: a concurrent dlopen never sees a short file
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
