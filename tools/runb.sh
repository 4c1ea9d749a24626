python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('evals/s',d['value'], 'ms',d['ms_per_step'], 'e2e',d['e2e']['value'], 'frac',d['roofline']['frac'])"
