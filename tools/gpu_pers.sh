#!/bin/bash
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in 0 1; do
  echo "PERSISTENT=$v: $(TSR_CONV_PERSISTENT=$v timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-170)"
done
TSR_CONV_PERSISTENT=1 timeout 100 python tools/bench_infer.py 1 512 2>&1 | tail -1
TSR_CONV_PERSISTENT=0 timeout 100 python tools/bench_infer.py 1 512 2>&1 | tail -1
for v in 0 1; do
  echo "B64 PERSISTENT=$v: $(TSR_CONV_PERSISTENT=$v timeout 120 python bench.py --batch 64 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-170)"
done
