# scratch command file for `gpurun -- bash tools/_scratch_run.sh` (overwritten per experiment)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 > gpurun_out/pytest_final2.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_final2.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 200 gpurun_out/bench_final2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1])
print('b16', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'b64', d['b64']['value'], 'infer', d['inference']['value'], d['inference']['e2e']['value'], 'esrgan', d['esrgan']['value'], 'launches', d['launches_per_step'], d['clocks'], 'x eager', d['gpu_eager_baseline']['speedup_over_best_variant'])
"
