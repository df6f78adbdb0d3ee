timeout 200 python tools/bench_programs.py 16 2>&1 | tail -5 | cut -c1-100
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "stage or oracle or golden or pair or residual or gan_step" 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 --only b64 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'b64', d['b64']['value'], 'launches', d['launches_per_step'])"
