"""Build ``libtsim_b200.so`` in-tree with nvcc for sm_100a (no JIT cache, no CPU variant)."""

from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "tsim_b200.cu")
HOST_SRC = os.path.join(HERE, "csrc", "host_pack.cpp")  # plain C++ (SIMD bit packing with run-time dispatch), g++
LIB = os.path.join(HERE, "libtsim_b200.so")
DEPS = [
    SRC,
    HOST_SRC,
    os.path.join(HERE, "csrc", "sampler_kernels.cuh"),
    os.path.join(HERE, "csrc", "zomega.cuh"),
    os.path.join(HERE, "csrc", "blob.h"),
    os.path.join(HERE, "csrc", "sliced_kernels.cuh"),
    os.path.join(HERE, "csrc", "noise_kernels.cuh"),
    os.path.join(HERE, "csrc", "postselect.cuh"),
    os.path.join(os.path.dirname(HERE), "include", "tsim_b200.h"),
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]
# `-split-compile 0` halves the build time but changes register allocation (the headline kernel spills 120 bytes with it):
# development builds only (TSIM_B200_FAST_BUILD=1), never the library that is measured
if os.environ.get("TSIM_B200_FAST_BUILD"):
    NVCC_FLAGS += ["-split-compile", "0"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libtsim_b200.so")
    gxx = shutil.which("g++") or "g++"
    host_obj = os.path.join(HERE, "csrc", "host_pack.o")
    res = subprocess.run([gxx, "-O3", "-fPIC", "-std=c++17", "-c", HOST_SRC, "-o", host_obj], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, SRC, host_obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
