#!/bin/bash
# Builds A/B variants of the library into variants/ (git-ignored, travels with gpurun):
#   usage: scripts/build_variants.sh name "-DFLAG=1 -DOTHER=2" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -maxrregcount=112 \
    -fmad=false -Xcompiler -fPIC -shared -Xptxas -v $flags -o variants/lib_$name.so \
    fgumi_b200/csrc/capi.cu fgumi_b200/csrc/host_tables.cpp fgumi_b200/csrc/host/caller_host.cpp fgumi_b200/csrc/host/bgzf.cpp 2> variants/$name.ptxas
  grep -A2 "Function properties for _ZN3fgb11vote_kernel" variants/$name.ptxas | tr '\n' ' '; echo " <- $name"
done
