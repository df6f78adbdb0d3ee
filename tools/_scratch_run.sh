mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 > gpurun_out/pytest_r2k.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2k.log | tail -20
timeout 600 python bench.py --steps 30 --warmup 5 --only hbm,b64,inference > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; tail -c 300 gpurun_out/bench_r2k.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2k.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','launches_per_step'):
    print(k, json.dumps(d.get(k))[:300])
print('b64', d['b64']['value'], 'inference', d['inference']['value'], d['inference']['e2e']['value'])
for h in d['hbm_kernels']: print(h['kernel'], round(h['us_per_step'],1), 'us', round(h['achieved_gbs']), 'GB/s', round(h['frac'],3))
PY
TIMELINE=gpurun_out/timeline_r2k.csv TOP=12 timeout 200 python tools/profile_step.py 16 2>&1 | tail -13 | cut -c1-160
