mkdir -p gpurun_out
for v in default s3c13; do
  if [ "$v" = default ]; then unset FGUMI_B200_LIB; else export FGUMI_B200_LIB=$PWD/variants/lib_$v.so; fi
  echo "=== $v"
  timeout 300 python scripts/depth_sweep.py 1000000 2 1,2,3,4,8,12,24,50,mixed2-20,zipf1-100 2>&1 | tail -10
done > gpurun_out/iter8_ab.log 2>&1
cat gpurun_out/iter8_ab.log
