"""Builds libgraphaudio_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m graphaudio_b200.build [--force]

The .so lands in graphaudio_b200/lib/ (git-ignored, but it travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libgraphaudio_cuda.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]

# (source, extra flags).  nodes.cu is compiled without FMA contraction: the biquad recursion, the cubic
# resampler polynomial and the automation curves must round like the reference's scalar float/double code.
SOURCES = [
    ("mac.cu", []),
    ("fft.cu", []),
    ("fft2.cu", []),
    ("fft_r16.cu", []),
    ("nodes.cu", ["--fmad=false"]),
    ("biquad.cu", ["--fmad=false"]),
    ("biquad_lanes.cu", ["--fmad=false"]),
    ("engine.cu", ["-Xcompiler", "-fvisibility=default"]),
]
HEADERS = ["gac_kernels.h", "fft2_core.cuh", "biquad_math.cuh", "engine_render.inl", "engine_stream.inl", os.path.join("..", "..", "include", "graphaudio_cuda.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs + [os.path.abspath(__file__)]):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
