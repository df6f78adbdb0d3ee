mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 > gpurun_out/pytest_r2m.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2m.log | tail -20
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; tail -c 300 gpurun_out/bench_r2m.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2m.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','launches_per_step','data_pipeline'):
    print(k, json.dumps(d.get(k))[:400])
print('b64', d['b64']['value'], 'inference', d['inference']['value'], 'esrgan', d['esrgan']['value'])
print(json.dumps(d['gpu_eager_baseline'])[:600])
PY
