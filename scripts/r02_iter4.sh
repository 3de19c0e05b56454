set -x
mkdir -p gpurun_out
timeout 300 python scripts/depth_sweep.py 1000000 2 1,2,3 > gpurun_out/depth_sweep_iter4.log 2>&1; tail -3 gpurun_out/depth_sweep_iter4.log
(timeout 600 python -m pytest tests/test_vote_parity.py -m gpu -x -q 2>&1 | tail -5)
for d in 24 100; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:vote_kernel_deep -s 1 -c 1 -o gpurun_out/r02_deepflat_d$d -f python scripts/profile_vote.py 400000 $d 3 > gpurun_out/ncu_deepflat_d$d.log 2>&1; tail -1 gpurun_out/ncu_deepflat_d$d.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vote_kernel_shallow -s 1 -c 1 -o gpurun_out/r02_shallow2_d1 -f python scripts/profile_vote.py 1000000 1 3 > gpurun_out/ncu_shallow2_d1.log 2>&1; tail -1 gpurun_out/ncu_shallow2_d1.log
