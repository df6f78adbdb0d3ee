mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/dp_check.py 16 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_r2f_n2.json 2> gpurun_out/bench_r2f_n2.err; tail -c 600 gpurun_out/bench_r2f_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2f_n2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','launches_per_step','dp_parity','b64'):
    print(k, json.dumps(d.get(k))[:600])
PY
