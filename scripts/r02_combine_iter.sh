# K2 / K3 / K4 iteration on the GPU box: parity tests for the combines, timings, optional ncu of the new kernels
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_combine_parity.py tests/test_caller_parity.py tests/test_filter.py tests/test_duplex_filter.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/combine_tests.log; cat gpurun_out/combine_tests.log
timeout 600 python scripts/bench_modes.py > gpurun_out/modes_new.jsonl 2>&1; cat gpurun_out/modes_new.jsonl
if [ -n "$NCU" ]; then
for k in $NCU; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_${k}_new -f python scripts/bench_modes.py 0.2 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
fi
