#!/bin/bash
# A/B of a library switch (env var named by $1, default TSR_PDL) on the GPU box: GPU tests, short bench runs with the
# switch off/on, per-program replay times, step profile.
VAR=${1:-TSR_PDL}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in 0 1; do
  ( env $VAR=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -3 ) > gpurun_out/bench_${VAR}$v.log
  ( env $VAR=$v timeout 200 python tools/bench_programs.py 16 2>&1 | tail -12 ) | tee gpurun_out/progs_${VAR}$v.log
done
( TOP=30 timeout 300 python tools/profile_step.py 16 2>&1 | tail -40 ) > gpurun_out/profile_${VAR}1.log
python - <<PY
import json
for v in (0,1):
    try:
        l=[x for x in open(f"gpurun_out/bench_${VAR}{v}.log") if x.startswith("{")][-1]
        d=json.loads(l); print("${VAR}",v,round(d["value"],1),"crops/s", round(d["ms_per_step"],3),"ms e2e",round(d["e2e"]["value"],1))
    except Exception as e: print("${VAR}",v,"failed",e)
PY
