#!/bin/bash
for d in 0 8 16 24; do
  echo "== DBG=$d"
  TSR_SPLIT_DBG=$d TSR_PDL=0 timeout 100 python tools/bench_programs.py 16 ops 2>&1 | grep -E "conv M=9216 N=64 K=9x64 bn=64 s=1 mode=0 stats=1 res=0|conv M=2304 N=512 K=9x256|conv M=576 N=512 K=9x512" | head -3
done
