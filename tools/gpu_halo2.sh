#!/bin/bash
TSR_CONV_HALO=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -6 | cut -c1-300
for v in 0 1; do
echo "== per-op HALO=$v"
TSR_CONV_HALO=$v TSR_PDL=0 timeout 100 python tools/bench_programs.py 16 ops 2>&1 | grep -E "conv M=9216 N=64 K=9x64|conv M=36864 N=256|conv M=147456|conv M=36864 N=64 K=9x64|conv M=36864 N=128" | head -8
done
