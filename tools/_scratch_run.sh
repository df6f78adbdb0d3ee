mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout=600 2>&1 > gpurun_out/pytest_r2l.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2l.log | tail -20; grep -E "stagewise summary|23 RRDB" gpurun_out/pytest_r2l.log | cut -c1-700
timeout 600 python bench.py --steps 30 --warmup 5 --only esrgan > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; tail -c 300 gpurun_out/bench_r2l.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2l.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','launches_per_step'):
    print(k, json.dumps(d.get(k))[:300])
print('esrgan', d['esrgan'].get('value'), d['esrgan'].get('ms_per_step'))
PY
TIMELINE=gpurun_out/timeline_r2l.csv TOP=2 timeout 200 python tools/profile_step.py 16 2>&1 | tail -3 | cut -c1-160
