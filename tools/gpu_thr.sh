#!/bin/bash
for v in 20 15 10; do
  echo "MIN_TILES_X10=$v: $(TSR_PERSIST_MIN_TILES_X10=$v timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-170)"
done
