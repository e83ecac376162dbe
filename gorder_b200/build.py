"""Build the CUDA engine in-tree: ``gorder_b200/libgorder_b200.so`` (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("GORDER_B200_LIB") or os.path.join(_HERE, "libgorder_b200.so")   # the override is for A/B builds
SOURCES = [os.path.join(_HERE, "csrc", "gorder_capi.cu")]
HEADERS = [
    os.path.join(_HERE, "csrc", "gorder_kernels.cuh"),
    os.path.join(_HERE, "csrc", "gorder_fast.cuh"),
    os.path.join(_HERE, "csrc", "gorder_ua_fast.cuh"),
    os.path.join(_HERE, "csrc", "gorder_spherical.cuh"),
    os.path.join(_HERE, "csrc", "gorder_xtc.inl"),
    os.path.join(_HERE, "csrc", "gorder_results.inl"),
    os.path.join(_HERE, "csrc", "gorder_multi.inl"),
    os.path.join(_HERE, "csrc", "gorder_topology.inl"),
    os.path.join(_HERE, "csrc", "gorder_engine.cuh"),
    os.path.join(_HERE, "csrc", "gorder_math.cuh"),
    os.path.join(_HERE, "..", "include", "gorder_b200.h"),
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA engine cannot be built")
    return p


def is_stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the engine if the shared library is missing or older than its sources."""
    if not force and not is_stale():
        return SO_PATH
    tmp = f"{SO_PATH}.tmp.{os.getpid()}"   # never leave a half-written library where a loader (or a snapshot) can see it
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
    env = dict(os.environ)
    # the image exports CC=/opt/gcc/bin/gcc; nvcc must use the system host compiler
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    os.replace(tmp, SO_PATH)
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
