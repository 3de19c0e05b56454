mkdir -p gpurun_out; rm -f gpurun_out/duplex_ab.log
for v in default "$@"; do
  if [ "$v" = default ]; then unset FGUMI_B200_LIB; else export FGUMI_B200_LIB=$PWD/variants/lib_$v.so; fi
  timeout 200 python scripts/duplex_ab.py 2000000 2>&1 | tail -1 >> gpurun_out/duplex_ab.log
done
cat gpurun_out/duplex_ab.log
