# K2 / K3 / K4 iteration on the GPU box: parity tests for the combines, timings, optional ncu of the new kernels
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_combine_parity.py tests/test_caller_parity.py tests/test_filter.py tests/test_duplex_filter.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/combine_tests.log; cat gpurun_out/combine_tests.log
timeout 600 python scripts/bench_modes.py > gpurun_out/modes_new.jsonl 2>&1; cat gpurun_out/modes_new.jsonl
if [ -n "$NCU" ]; then
for k in $NCU; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_${k}_new -f python scripts/bench_modes.py 0.2 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
fi
FGB_SUBMIT_TRACE=1 timeout 300 python scripts/bench_records.py 200000 16 > gpurun_out/records_trace.log 2>&1; tail -60 gpurun_out/records_trace.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/records_launches.csv python scripts/bench_records.py 200000 16 > gpurun_out/records_under_ncu.log 2>&1; python - <<'P'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/records_launches.csv') if l.startswith('"')))
h = rows[0]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        agg[r[ki].split('(')[0]][0] += 1; agg[r[ki].split('(')[0]][1] += float(r[vi].replace(',', ''))
    except Exception:
        pass
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} launches {n:4d}  total {t/1e6:9.3f} ms  mean {t/n/1e3:9.1f} us")
P
