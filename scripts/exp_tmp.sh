python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
print({k:(round(v['ms_per_step']/v['launches_per_step']*1000,1), v['launches_per_step']) for k,v in d['kernel_breakdown'].items()})"
