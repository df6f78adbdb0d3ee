mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 > gpurun_out/pytest_r2o.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2o.log | tail -10
timeout 300 python tools/bench_infer.py 2>&1 | tail -1
timeout 300 python tools/profile_infer.py 2>&1 | grep -E "im2row|sum of" 
timeout 900 python bench.py > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; tail -c 300 gpurun_out/bench_r2o.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2o.json').read().strip().splitlines()[-1])
print('b16', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'b64', d['b64']['value'], 'infer', d['inference']['value'], d['inference']['e2e']['value'], 'esrgan', d['esrgan']['value'], 'launches', d['launches_per_step'])
print(json.dumps(d['roofline']['dominant_kernel']))
"
