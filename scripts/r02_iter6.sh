set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/iter6_pytest.log; cat gpurun_out/iter6_pytest.log
timeout 400 python scripts/bench_modes.py 2>&1 | cut -c1-700 > gpurun_out/iter6_modes.jsonl; cat gpurun_out/iter6_modes.jsonl
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/iter6_memcheck.log 2>&1; tail -4 gpurun_out/iter6_memcheck.log
FGB_BIND_NUMA=1 timeout 300 python scripts/bench_records.py 200000 16 2>&1 | grep "^rep\|^{" | tail -5 | cut -c1-400
