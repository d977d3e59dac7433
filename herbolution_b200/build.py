"""Build libherbgpu.so (sm_100a CUDA, in-tree).

    python -m herbolution_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  -fmad=false is global: the noise path must not
contract mul+add (only the six explicit __fmaf_rn sites fuse, see csrc/noise.cuh).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "herbolution_b200", "csrc")
LIB = os.path.join(ROOT, "herbolution_b200", "libherbgpu.so")

CUDA_SOURCES = ["gen.cu", "mesh.cu", "api.cu", "world.cu", "codec.cu", "expand.cu", "multi.cu"]
CUDA_DEPS = CUDA_SOURCES + ["common.cuh", "noise.cuh", "host_common.cuh", os.path.join("..", "..", "include", "herbgpu.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libherbgpu.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    deps = [os.path.join(CSRC, d) for d in CUDA_DEPS] + [os.path.abspath(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in CUDA_SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(ROOT, "herbolution_b200", "csrc", "ptxas.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libherbgpu.so")
    return LIB


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_cuda(force=force, verbose=True))
