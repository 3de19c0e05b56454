set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/iter5_tests.log; cat gpurun_out/iter5_tests.log
timeout 300 python scripts/depth_sweep.py 1000000 2 1,2,3,4,6,8,12,20,24,32,50,100,mixed2-20,zipf1-100 > gpurun_out/depth_sweep_iter5.log 2>&1; tail -14 gpurun_out/depth_sweep_iter5.log
FGB_BIND_NUMA=1 timeout 300 python scripts/bench_records.py 200000 16 > gpurun_out/records_iter5.log 2>&1; grep "^rep\|fgb_caller" gpurun_out/records_iter5.log | tail -8; tail -1 gpurun_out/records_iter5.log | cut -c1-330
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r02_memcheck.log 2>&1; tail -4 gpurun_out/r02_memcheck.log
