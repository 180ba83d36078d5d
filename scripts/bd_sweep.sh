#!/bin/bash
# per-launch kernel times (us) of the F2 step at several batch sizes (latency- vs bandwidth-bound check)
for bd in "$@"; do
  timeout 300 python bench.py --no-cpu-baseline --only-f2 --bd $bd --steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('bd', $bd, 'ms_per_step', d['ms_per_step'], {k:(round(v['ms']/v['launches']*1000,1)) for k,v in d['kernel_breakdown'].items()})"
done
