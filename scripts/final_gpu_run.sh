mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
bash scripts/run_bench_gpu.sh > gpurun_out/run_bench.log 2>&1; tail -4 gpurun_out/run_bench.log | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vote_kernel -s 2 -c 1 -o gpurun_out/vote_r01k -f python scripts/profile_vote.py 1000000 8 4 > gpurun_out/ncu_r01k.log 2>&1; tail -2 gpurun_out/ncu_r01k.log | cut -c1-200
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/memcheck.log 2>&1; tail -3 gpurun_out/memcheck.log
