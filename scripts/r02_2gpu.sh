set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_2gpu_line.json 2> gpurun_out/r02_bench_2gpu.err; echo rc=$?; tail -5 gpurun_out/r02_bench_2gpu.err; cut -c1-300 gpurun_out/r02_bench_2gpu_line.json
(timeout 600 python -m pytest tests/test_vote_parity.py tests/test_combine_parity.py -m gpu -x -q 2>&1 | tail -3)
timeout 300 python scripts/depth_sweep.py 1000000 2 1,2,3,4,8,24,50,100 2>&1 | tail -8
