#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
timeout 100 python tools/bench_infer.py 1 512 2>&1 | tail -1
echo "B64: $(timeout 120 python bench.py --batch 64 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-170)"
echo "HALO=1: $(TSR_CONV_HALO=1 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-170)"
TSR_CONV_HALO=1 timeout 100 python tools/bench_infer.py 1 512 2>&1 | tail -1
