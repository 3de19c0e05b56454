set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; echo rc=$?
tail -3 gpurun_out/bench_main.err
cat gpurun_out/bench_main.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "fgb_timed/" -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 5 --warmup 3 --cpu-units 0 > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?
tail -5 gpurun_out/r01_launches.csv
