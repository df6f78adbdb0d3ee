mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 --only none > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; tail -c 300 gpurun_out/bench_r2e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2e.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','launches_per_step'):
    print(k, json.dumps(d.get(k))[:700])
PY
TIMELINE=gpurun_out/timeline_r2e.csv TOP=3 timeout 200 python tools/profile_step.py 16 2>&1 | tail -4 | cut -c1-160
