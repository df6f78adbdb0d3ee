timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -x -k "generator or golden or gan_step or pretrain or stage or cli or packed" 2>&1 | tail -2
timeout 600 python bench.py --only b64 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b16', round(d['value']), d['ms_per_step'], 'b64', round(d['b64']['value']), d['b64']['ms_per_step'])"
