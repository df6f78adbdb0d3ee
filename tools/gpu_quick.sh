#!/bin/bash
timeout 200 python -m pytest tests/test_modules_gpu.py -m gpu -x -q -k "adam or packed or gan_step" 2>&1 | tail -2
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
TIMELINE=gpurun_out/timeline_b16_latest.csv TOP=1 timeout 120 python tools/profile_step.py 16 2>&1 | tail -2 | cut -c1-160
grep -E "adam|pack" gpurun_out/timeline_b16_latest.csv | cut -c1-80
