export TSR_OCCUPANCY=own TSR_PDL_BIG_BARRIER=0
TOP=2 timeout 100 python tools/profile_step.py 64 2>&1 | tail -3 | cut -c1-160 > gpurun_out/exp_bigbarrier.log; cat gpurun_out/exp_bigbarrier.log
if grep -q "GPU span [0-9]\.[0-9]* ms" gpurun_out/exp_bigbarrier.log; then
timeout 200 python bench.py --only b64 --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('own+nopdl-big: b16', round(d['value']), d['ms_per_step'], 'b64', round(d['b64']['value']), d['b64']['ms_per_step'], d['b64']['launches_per_step'])"
fi
