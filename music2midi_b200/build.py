"""Builds libm2m_b200.so in-tree with nvcc for sm_100a.

The library has a plain C ABI (include/m2m_b200.h) and is loaded with ctypes, so no torch
extension machinery is involved.  The .so is git-ignored but travels to the GPU box.

    python -m music2midi_b200.build [--force] [-v]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libm2m_b200.so")
SOURCES = ["api.cu", "notes.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libm2m_b200.so for sm_100a)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "m2m_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d) and not d.endswith(".o"))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc(), *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
               "-Xptxas", "-v" if verbose else "-warn-spills", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [_nvcc(), *ARCH, "-shared", "-cudart", "static", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


NOTES_LIB = os.path.join(HERE, "libm2m_notes.so")


def build_notes() -> str:
    """Host-only build of the token -> notes state machine (g++, no CUDA): libm2m_notes.so."""
    cxx = shutil.which("g++") or shutil.which("c++")
    if not cxx:
        raise RuntimeError("g++ not found (needed to build libm2m_notes.so)")
    cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-DM2M_NOTES_STANDALONE", os.path.join(CSRC, "notes.cpp"), "-o", NOTES_LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed on notes.cpp")
    return NOTES_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
