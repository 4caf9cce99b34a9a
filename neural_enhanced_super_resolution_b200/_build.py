"""Builds ``libnesr_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The shared library is pure CUDA runtime + driver entry points (no torch, no Python): the drop-in
boundary is the C ABI in ``include/nesr_b200.h``.  Called by ``__graft_entry__.build()``; can also be
run as ``python -m neural_enhanced_super_resolution_b200._build``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libnesr_b200.so")
SOURCES = ["conv3x3_fold.cu", "conv3x3_body.cu", "conv3x3_trunk.cu", "conv3x3_simt.cu", "pixel_io.cu", "stencil.cu", "sharpen_mma.cu", "preprocess.cu", "engine.cu"]
HEADERS = ["ptx.cuh", "layout.h", "epilogue.cuh", "fold_roles.cuh", "kernels.h", "lab_tables.inc", os.path.join("..", "..", "include", "nesr_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# Timing-experiment variant (device printf, per-role cycle accounting, the NESR_B200_DEBUG_FLAGS switches): built as a
# SEPARATE library, libnesr_b200_prof.so, selected at run time with NESR_B200_LIB=<path>; the product library never contains it.
PROF_LIB = os.path.join(PKG, "libnesr_b200_prof.so")


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libnesr_b200.so cannot be built")
    return nvcc


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, prof: bool = False) -> str:
    nvcc = _nvcc()
    OBJ = os.path.join(PKG, "build_prof" if prof else "build")
    LIB = PROF_LIB if prof else globals()["LIB"]
    FLAGS = globals()["FLAGS"] + (["-DNESR_PROF=1"] if prof else [])
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in HEADERS]
    stamp = os.path.join(OBJ, "flags.txt")                    # a different flag set rebuilds everything
    flags_now = " ".join(ARCH + FLAGS)
    if not os.path.exists(stamp) or open(stamp).read() != flags_now:
        force = True
        with open(stamp, "w") as fh:
            fh.write(flags_now)
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *ARCH, *FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, prof="--prof" in sys.argv))
