mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/dp_check.py 16 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_r2h_n2.json 2> gpurun_out/bench_r2h_n2.err; tail -c 600 gpurun_out/bench_r2h_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2h_n2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','dp_parity','b64'):
    print(k, json.dumps(d.get(k))[:300])
PY
TIMELINE=gpurun_out/timeline_r2h_n2.csv TOP=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 tools/profile_step.py 16 2>&1 | tail -3 | cut -c1-170
