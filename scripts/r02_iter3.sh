# iteration 3: flat deep kernel + shared-memory pair table: parity, depth sweep, zipf leg; records path slots / NUMA
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_vote_parity.py tests/test_golden.py tests/test_full_size.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/iter3_tests.log; cat gpurun_out/iter3_tests.log
timeout 600 python scripts/depth_sweep.py 1000000 2 1,2,3,4,8,24,32,50,100,mixed2-20,zipf1-100 > gpurun_out/depth_sweep_iter3.log 2>&1; tail -12 gpurun_out/depth_sweep_iter3.log
timeout 300 python scripts/bench_records.py 200000 16 > gpurun_out/records_slots2.log 2>&1; tail -1 gpurun_out/records_slots2.log | cut -c1-330
FGB_SUBMIT_SLOTS=4 timeout 300 python scripts/bench_records.py 200000 16 > gpurun_out/records_slots4.log 2>&1; tail -1 gpurun_out/records_slots4.log | cut -c1-330
FGB_BIND_NUMA=1 FGB_SUBMIT_SLOTS=4 FGB_SUBMIT_TRACE=1 timeout 300 python scripts/bench_records.py 200000 16 > gpurun_out/records_slots4_numa.log 2>&1; tail -1 gpurun_out/records_slots4_numa.log | cut -c1-330; sed -n 101,150p gpurun_out/records_slots4_numa.log
