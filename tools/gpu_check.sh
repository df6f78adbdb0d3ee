#!/bin/bash
# Standard GPU round-trip: every step under its own short timeout so that a hang costs seconds, not the call's limit.
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
timeout 120 python tools/bench_programs.py 16 2>&1 | tail -6 | cut -c1-90
TIMELINE=gpurun_out/timeline_b16_latest.csv TOP=1 timeout 120 python tools/profile_step.py 16 2>&1 | tail -2 | cut -c1-160
