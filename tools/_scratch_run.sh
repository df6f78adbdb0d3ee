mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 > gpurun_out/pytest_r2i.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2i.log | tail -20
timeout 200 python tools/bench_programs.py 16 2>&1 | tail -6 | cut -c1-120
timeout 600 python bench.py --steps 30 --warmup 5 --only none > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -c 300 gpurun_out/bench_r2i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2i.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','launches_per_step'):
    print(k, json.dumps(d.get(k))[:300])
PY
