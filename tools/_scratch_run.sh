mkdir -p gpurun_out
for h in 0 1; do
TSR_CONV_HALO=$h timeout 600 python bench.py --steps 30 --warmup 5 --only b64,inference > gpurun_out/bench_halo$h.json 2> gpurun_out/bench_halo$h.err; tail -c 300 gpurun_out/bench_halo$h.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_halo$h.json').read().strip().splitlines()[-1])
print('HALO=$h', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'b64', round(d['b64']['value'],1), 'inference', round(d['inference']['value'],1))
PY
done
TSR_CONV_HALO=1 timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "oracle or stage or golden or psnr" 2>&1 | tail -3
