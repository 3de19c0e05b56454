#!/bin/bash
# GPU box: time each variants/lib_<name>.so on the depth sweep subset and the headline batch.
mkdir -p gpurun_out
for v in "$@"; do
  export FGUMI_B200_LIB=$PWD/variants/lib_$v.so
  echo "=== $v" >> gpurun_out/ab.log
  timeout 300 python scripts/depth_sweep.py 1000000 2 ${AB_SPECS:-2,3,4,8,mixed2-20,zipf1-100} 2>&1 | tail -8 >> gpurun_out/ab.log
  timeout 200 python scripts/profile_vote.py 10000000 8 6 2>&1 | tail -2 | head -1 >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
