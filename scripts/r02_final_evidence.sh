# Round-2 final evidence on the GPU box (one B200): tests, bench lines, launch list, ncu captures, sweeps, record level, file level.
set -x
mkdir -p gpurun_out
bash scripts/sysinfo.sh > gpurun_out/r02_gpu_box_sysinfo.txt 2>&1
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err; echo rc=$?; tail -3 gpurun_out/r02_bench.err; cut -c1-300 gpurun_out/r02_bench_line.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference_line.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "fgb_timed/" -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --cpu-units 0 --no-modes > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?; tail -3 gpurun_out/r02_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"vote_kernel$" -s 2 -c 1 -o gpurun_out/vote_r02 -f python scripts/profile_vote.py 1000000 8 4 > gpurun_out/ncu_vote_r02.log 2>&1; tail -2 gpurun_out/ncu_vote_r02.log | cut -c1-200
timeout 600 python scripts/depth_sweep.py 1000000 2 1,2,3,4,6,8,12,20,24,32,50,100,mixed2-20,zipf1-100 > gpurun_out/r02_depth_sweep.log 2>&1; tail -14 gpurun_out/r02_depth_sweep.log
timeout 600 python scripts/bench_modes.py > gpurun_out/r02_modes.jsonl 2>&1; cat gpurun_out/r02_modes.jsonl
FGB_BIND_NUMA=1 FGB_SUBMIT_TRACE=1 timeout 300 python scripts/bench_records.py 200000 16 > gpurun_out/r02_record_level_phases.log 2>&1; grep "^rep\|fgb_caller\|numa" gpurun_out/r02_record_level_phases.log | tail -10
timeout 300 python scripts/file_level_run.py > gpurun_out/r02_file_level_run.log 2>&1; tail -6 gpurun_out/r02_file_level_run.log
for k in duplex_combine_words_kernel codec_combine_words_kernel filter_simplex_words_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_$k -f python scripts/bench_modes.py 0.2 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
for k in unpack_records_kernel assemble_simplex_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/r02_$k -f python scripts/bench_records.py 200000 16 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r02_compute_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02_compute_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/r02_compute_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r02_compute_sanitizer_racecheck.log
