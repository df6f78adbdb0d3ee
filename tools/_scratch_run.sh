timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -x -k "gan_step or pretrain or stage or esrgan_generator or golden or trainer" 2>&1 | tail -2
timeout 400 python bench.py --only esrgan --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b16', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'esrgan', round(d['esrgan']['value']))"
