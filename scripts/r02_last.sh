set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 420 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err; echo rc=$?; wc -l gpurun_out/r02_bench_line.json; cut -c1-200 gpurun_out/r02_bench_line.json
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_reference_line.json
