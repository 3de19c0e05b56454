# iteration 2: combine + filter tests, records path after the K5-upload / chunk changes, ncu of the shallow and deep vote kernels
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_combine_parity.py tests/test_caller_parity.py tests/test_filter.py tests/test_records_input.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/iter2_tests.log; cat gpurun_out/iter2_tests.log
timeout 600 python scripts/bench_modes.py > gpurun_out/modes_iter2.jsonl 2>&1; cat gpurun_out/modes_iter2.jsonl
FGB_SUBMIT_TRACE=1 timeout 300 python scripts/bench_records.py 200000 16 > gpurun_out/records_trace2.log 2>&1; grep -v "chunk" gpurun_out/records_trace2.log | tail -30; sed -n 20,60p gpurun_out/records_trace2.log
for d in 1 2 3; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:vote_kernel_shallow -s 1 -c 1 -o gpurun_out/r02_shallow_d$d -f python scripts/profile_vote.py 1000000 $d 3 > gpurun_out/ncu_shallow_d$d.log 2>&1; tail -1 gpurun_out/ncu_shallow_d$d.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vote_kernel_deep -s 1 -c 1 -o gpurun_out/r02_deep_d100 -f python scripts/profile_vote.py 400000 100 3 > gpurun_out/ncu_deep_d100.log 2>&1; tail -1 gpurun_out/ncu_deep_d100.log
