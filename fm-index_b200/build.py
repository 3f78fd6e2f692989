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
    """Compile fm-index_b200/csrc into fm-index_b200/libfmx_b200.so with nvcc for sm_100a: one object per source,
    compiled side by side (objects under csrc/_obj, git-ignored), then one link."""
    if not force and not needs_build():
        return LIB
    env = dict(os.environ)
    env.pop("CC", None)   # a broken gcc wrapper in CC/CXX must not leak into nvcc's host compile
    env.pop("CXX", None)
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    newest_header = max(os.path.getmtime(os.path.join(CSRC, d)) for d in DEPS if d not in SOURCES)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(os.path.join(CSRC, src)), newest_header)
        if stale:
            procs.append((src, subprocess.Popen([_nvcc()] + flags + ["-c", "-o", obj, src], cwd=CSRC, env=env,
                                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = ""
    for src, pr in procs:
        out = pr.communicate()[0]
        log += out
        if pr.returncode != 0:
            for _, other in procs:
                if other.poll() is None:
                    other.kill()
            raise RuntimeError(f"nvcc failed on {src}:\n" + out)
    objs = [os.path.join(objdir, os.path.splitext(src)[0] + ".o") for src in SOURCES]
    res = subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lgomp"],
                         cwd=CSRC, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(log + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
