timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "gather_out or gan_step or discriminator_two or frozen" 2>&1 | tail -2
timeout 400 python bench.py --only none --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b16', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))"
