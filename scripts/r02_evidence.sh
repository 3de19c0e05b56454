# Round-2 evidence run on the GPU box: tests, bench lines, sweeps, ncu captures of the secondary kernels.
set -x
mkdir -p gpurun_out
bash scripts/sysinfo.sh > gpurun_out/r02_gpu_box_sysinfo.txt 2>&1
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err; echo rc=$?; tail -3 gpurun_out/r02_bench.err; cut -c1-600 gpurun_out/r02_bench_line.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_line.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_reference_line.json
timeout 600 python scripts/depth_sweep.py 1000000 2 1,2,3,4,6,8,12,20,24,32,50,100,mixed2-20,zipf1-100 > gpurun_out/r02_depth_sweep.log 2>&1; tail -20 gpurun_out/r02_depth_sweep.log
timeout 600 python scripts/bench_modes.py > gpurun_out/r02_modes.jsonl 2>&1; cat gpurun_out/r02_modes.jsonl
timeout 300 python scripts/bench_records.py 200000 16 > gpurun_out/r02_record_level_phases.log 2>&1; tail -30 gpurun_out/r02_record_level_phases.log
timeout 300 python scripts/file_level_run.py > gpurun_out/r02_file_level_run.log 2>&1; tail -8 gpurun_out/r02_file_level_run.log
for k in duplex_combine_kernel codec_combine_kernel filter_simplex_kernel unpack_bam4_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_$k -f python scripts/bench_modes.py 0.2 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
