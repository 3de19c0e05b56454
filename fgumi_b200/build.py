"""In-tree build of libfgumi_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfgumi_b200.so")
SOURCES = ["capi.cu", "host_tables.cpp", "host/caller_host.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-maxrregcount=112",      # vote_kernel: 2 CTAs x 288 threads x 112 regs = 64512 <= 65536 per SM
    "-fmad=false",            # no FMA contraction anywhere near the f64 vote (DESIGN.md numerics)
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    deps = [os.path.join(dp, f) for dp, _, fs in os.walk(CSRC) for f in fs]
    deps.append(os.path.join(HERE, "..", "include", "fgumi_b200.h"))
    if not force and _newer(OUT, deps):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libfgumi_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
