# e2e sweep over submit pipeline depth / chunk size (GPU box)
for cfg in "2 96" "3 64" "4 32" "3 32" "2 48" "4 64"; do
  set -- $cfg
  FGB_SUBMIT_SLOTS=$1 FGB_SUBMIT_CHUNK_MB=$2 timeout 300 python bench.py --cpu-units 0 --units 1000000 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 slots $2 MB:', round(d['e2e']['value']/1e6,2), 'M/s pack8;', round(d['e2e']['two_column']['value']/1e6,2), 'M/s two-column')"
done
