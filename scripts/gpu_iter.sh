# One optimisation iteration on the GPU box: kernel parity tests, timing, optional ncu capture.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vote_parity.py tests/test_combine_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python scripts/profile_vote.py 10000000 8 8 2>&1 | tail -3
if [ -n "$NCU_TAG" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:vote_kernel -s 2 -c 1 -o gpurun_out/vote_$NCU_TAG -f python scripts/profile_vote.py 1000000 8 4 > gpurun_out/ncu_$NCU_TAG.log 2>&1
  tail -2 gpurun_out/ncu_$NCU_TAG.log
fi
