#!/bin/bash
mkdir -p gpurun_out
B=${1:-16}
for v in 0 1; do
  ( TSR_PDL=$v timeout 200 python tools/bench_programs.py $B 2>&1 | tail -12 ) | tee gpurun_out/progs_pdl$v.log
done
( TSR_WGRAD_BRANCH=0 timeout 200 python tools/bench_programs.py $B 2>&1 | tail -12 ) | tee gpurun_out/progs_nobranch.log
