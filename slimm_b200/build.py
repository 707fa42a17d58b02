"""Builds the in-tree native library ``slimm_b200/libslimm_gpu.so`` (CUDA kernels for sm_100a + C ABI).

    python -m slimm_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with gpurun.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libslimm_gpu.so")
CLI = os.path.join(HERE, "bin", "slimm")            # drop-in command line (csrc/frontend), links against LIB
FRONTEND = os.path.join(CSRC, "frontend")
CXX_FLAGS = ["-O2", "-std=c++17", "-Wall", "-ffp-contract=off", "-fno-fast-math", "-pthread"]
SOURCES = ["slimm_gpu.cu", "profile_host.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-Xptxas", "-v", "--shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, s))] + \
           [os.path.join(HERE, "..", "include", "slimm_gpu.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _cli_stale() -> bool:
    if not os.path.exists(CLI):
        return True
    t = os.path.getmtime(CLI)
    deps = [os.path.join(FRONTEND, s) for s in os.listdir(FRONTEND)] + [os.path.join(HERE, "..", "include", "slimm_gpu.h"), LIB,
                                                                         os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_cli(force: bool = False) -> str:
    """g++ -> slimm_b200/bin/slimm: the drop-in `slimm` command line (host decode pipeline, .sldb reader, TSV writers)
    on top of libslimm_gpu.so; found at run time through an $ORIGIN-relative rpath."""
    if not force and not _cli_stale():
        return CLI
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx] + CXX_FLAGS + ["-o", CLI, os.path.join(FRONTEND, "slimm_main.cpp"), "-L" + HERE, "-lslimm_gpu",
                               "-Wl,-rpath,$ORIGIN/..", "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building slimm_b200/bin/slimm")
    return CLI


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        build_cli(force)
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("SLIMM_NVCC_EXTRA", "").split() + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libslimm_gpu.so")
    log = os.path.join(HERE, "ptxas.log")   # git-ignored: registers / spills / shared memory per kernel of the last build
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    build_cli(True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
