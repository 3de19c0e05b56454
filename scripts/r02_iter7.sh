set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_vote_parity.py tests/test_full_size.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/iter7_pytest.log; cat gpurun_out/iter7_pytest.log
timeout 300 python scripts/depth_sweep.py 1000000 2 1,2,3,4,mixed2-20,zipf1-100 2>&1 | tail -6 > gpurun_out/iter7_sweep_defer.log; cat gpurun_out/iter7_sweep_defer.log
FGUMI_B200_LIB=$PWD/variants/lib_nodefer.so timeout 300 python scripts/depth_sweep.py 1000000 2 1,2,3,4,mixed2-20,zipf1-100 2>&1 | tail -6 > gpurun_out/iter7_sweep_nodefer.log; cat gpurun_out/iter7_sweep_nodefer.log
timeout 200 python scripts/duplex_ab.py 5000000 2>&1 | tail -1
