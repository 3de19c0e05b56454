# The essential round-end evidence in the order of importance (each step bounded): GPU tests, bench line, reference line, launch list.
set -x
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 420 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err; echo rc=$?; cut -c1-300 gpurun_out/r02_bench_line.json
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference_line.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "fgb_timed/" -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --cpu-units 0 --no-modes > gpurun_out/bench_under_ncu.log 2>&1; tail -3 gpurun_out/r02_launches.csv
timeout 300 python scripts/bench_modes.py > gpurun_out/r02_modes.jsonl 2>&1; cut -c1-400 gpurun_out/r02_modes.jsonl
timeout 200 python scripts/depth_sweep.py 1000000 2 1,2,3,4,6,8,12,20,24,32,50,100,mixed2-20,zipf1-100 > gpurun_out/r02_depth_sweep.log 2>&1; tail -14 gpurun_out/r02_depth_sweep.log
FGB_BIND_NUMA=1 FGB_SUBMIT_TRACE=1 timeout 200 python scripts/bench_records.py 200000 16 > gpurun_out/r02_record_level_phases.log 2>&1; grep "^rep\|fgb_caller\|numa" gpurun_out/r02_record_level_phases.log | tail -10
timeout 200 python scripts/file_level_run.py > gpurun_out/r02_file_level_run.log 2>&1; tail -4 gpurun_out/r02_file_level_run.log
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r02_compute_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02_compute_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/r02_compute_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r02_compute_sanitizer_racecheck.log
for k in duplex_combine_words_kernel filter_simplex_words_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_$k -f python scripts/bench_modes.py 0.2 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
for k in unpack_records_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/r02_$k -f python scripts/bench_records.py 200000 16 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
