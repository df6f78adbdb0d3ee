timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 > gpurun_out/pytest_r2u.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2u.log | tail -5
timeout 300 python tools/bench_infer.py 2>&1 | tail -1 | cut -c60-170
timeout 300 python tools/profile_infer.py 2>&1 | sed -n 3,6p | cut -c1-100
timeout 600 python bench.py --only b64 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b16', round(d['value']), d['ms_per_step'], 'b64', round(d['b64']['value']), d['b64']['ms_per_step'])"
