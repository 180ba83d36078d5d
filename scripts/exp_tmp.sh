python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tests/diag_chain.py 9472 2>&1 | grep -B1 -A12 "CTA 74" | head -14
python bench.py --no-cpu-baseline --only-f2 --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step', d['ms_per_step'], {k:(round(v['ms']/v['launches']*1000,1)) for k,v in d['kernel_breakdown'].items()})"
