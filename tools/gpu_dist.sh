#!/bin/bash
N=${1:-2}
echo "== N=$N"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | grep -E "^\{|Error|error|Traceback|timed out" | cut -c1-330 | tail -5
echo "exit: $?"
