#!/bin/bash
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 8 16; do
  echo "SPLITK_MIN_ITERS=$v: $(TSR_SPLITK_MIN_ITERS=$v timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-170)"
  TSR_SPLITK_MIN_ITERS=$v timeout 120 python tools/bench_programs.py 16 2>&1 | tail -5 | cut -c1-60
done
