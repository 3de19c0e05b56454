mkdir -p gpurun_out; rm -f gpurun_out/iter9.log
for v in default dc64 dc128 dc256; do
  if [ "$v" = default ]; then unset FGUMI_B200_LIB; else export FGUMI_B200_LIB=$PWD/variants/lib_$v.so; fi
  timeout 200 python scripts/duplex_ab.py 5000000 2>&1 | tail -1 >> gpurun_out/iter9.log
  (timeout 300 python -m pytest tests/test_combine_parity.py -m gpu -x -q -k duplex 2>&1 | tail -2) >> gpurun_out/iter9.log
done
cat gpurun_out/iter9.log
