# Duplex epilogue: parity, sanitizer, timings (one B200).
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_combine_parity.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/fused_pytest_combine.log; cat gpurun_out/fused_pytest_combine.log
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/fused_memcheck.log 2>&1; tail -4 gpurun_out/fused_memcheck.log
timeout 400 python scripts/bench_modes.py > gpurun_out/fused_modes.jsonl 2>&1; cut -c1-1500 gpurun_out/fused_modes.jsonl
(timeout 900 python -m pytest tests/test_full_size.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/fused_pytest_full.log; cat gpurun_out/fused_pytest_full.log
timeout 400 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/fused_racecheck.log 2>&1; tail -4 gpurun_out/fused_racecheck.log
