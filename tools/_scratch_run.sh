timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -x -k "esrgan or dense" 2>&1 | tail -3
timeout 600 python bench.py --only esrgan --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('esrgan', d['esrgan']['value'], d['esrgan']['ms_per_step'])"
