#!/bin/bash
# Round-2 GPU round trip: full GPU test-suite (no -x: every failure is wanted from one call), smoke, bench.
# Every step under its own timeout so that a hang costs seconds, not the call's limit.
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout=600 2>&1 > gpurun_out/pytest_${TAG}.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_${TAG}.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -3 gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 1500 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
