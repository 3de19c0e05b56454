mkdir -p gpurun_out; rm -f gpurun_out/iter10.log
for v in default cc128 cc256; do
  if [ "$v" = default ]; then unset FGUMI_B200_LIB; else export FGUMI_B200_LIB=$PWD/variants/lib_$v.so; fi
  echo "=== $v" >> gpurun_out/iter10.log
  timeout 300 python scripts/bench_modes.py 2>&1 | cut -c1-900 >> gpurun_out/iter10.log
  (timeout 400 python -m pytest tests/test_combine_parity.py tests/test_filter.py tests/test_duplex_filter.py -m gpu -x -q 2>&1 | tail -2) >> gpurun_out/iter10.log
done
cat gpurun_out/iter10.log
