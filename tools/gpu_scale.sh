#!/bin/bash
# Weak-scaling run on one box: bench.py at N = 1, 2, 4, 8 back to back (the driver's SCALE recipe) + dp check at N.
N=${1:-8}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 30 --warmup 5 --only b64 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+n)) bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1])
    print($n, 'crops/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'b64', round(d['b64']['value'],1), 'dp_parity', d.get('dp_parity'))
except Exception as e:
    print($n, 'failed', e); print(open('gpurun_out/scale_n$n.err').read()[-1500:])
PY
done
