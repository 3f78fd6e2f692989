"""Build recipe for the CUDA library (in-tree, sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfmx_b200.so")
SOURCES = ["fmx_api.cu", "fmx_group.cu", "gpu_sa.cu", "gpu_build.cu", "builder.cpp"]
DEPS = SOURCES + ["kernels.cuh", "phased.cuh", "scan3.cuh", "fmx_layout.h", "builder.h", "sais.hpp", os.path.join("..", "..", "include", "fmx.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fopenmp,-O3,-Wall",
    "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the fmx CUDA library cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile fm-index_b200/csrc into fm-index_b200/libfmx_b200.so with nvcc for sm_100a."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + ["-lgomp"]
    env = dict(os.environ)
    env.pop("CC", None)   # a broken gcc wrapper in CC/CXX must not leak into nvcc's host compile
    env.pop("CXX", None)
    res = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
