mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "kernel or large_image or pipelined or generator or golden or eval_mode" 2>&1 | tail -3
timeout 600 python tools/microbench_wgrad.py 2>/dev/null | tee gpurun_out/microbench_wgrad2.log
timeout 300 python tools/bench_infer.py 2>&1 | tail -1
timeout 900 python bench.py --only b64,inference > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; tail -c 300 gpurun_out/bench_r2p.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2p.json').read().strip().splitlines()[-1])
print('b16', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'b64', d['b64']['value'], 'infer', d['inference']['value'], d['inference']['e2e'], 'launches', d['launches_per_step'])
"
