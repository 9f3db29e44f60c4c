"""Builds libmsort.so in-tree with nvcc for sm_100a (B200).  No other target, no JIT cache."""
from __future__ import annotations

import os
import shutil
import subprocess

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = ["msort_kernels.cu", "msort_policy.cu", "msort_ppo.cu", "msort_api.cu"]
HEADERS = ["msort_device.cuh", "msort_umma.cuh", "msort_launch.h", os.path.join("..", "..", "include", "msort.h")]
LIB = os.path.join(_CSRC, "libmsort.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v", "--use_fast_math=false"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(p):
        raise RuntimeError("nvcc not found; libmsort.so cannot be built (there is no CPU fallback)")
    return p


def needs_build() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(_CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [nvcc_path()] + flags + ["-o", LIB] + [os.path.join(_CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, cwd=_CSRC, capture_output=True, text=True)
    log = os.path.join(_CSRC, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}); see {log}\n{res.stderr[-4000:]}")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
