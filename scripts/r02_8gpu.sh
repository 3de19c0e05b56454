# bench.py on N GPUs of one box under torchrun (N = $1, default 8): whole-job values, the config-5 Zipf split over N ranks
N=${1:-8}
mkdir -p gpurun_out
timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu_line.json 2> gpurun_out/r02_bench_${N}gpu.err; echo rc=$?; tail -3 gpurun_out/r02_bench_${N}gpu.err; cut -c1-400 gpurun_out/r02_bench_${N}gpu_line.json
