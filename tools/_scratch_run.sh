echo "A carveout=max, api occupancy"; timeout 200 python bench.py --only none --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
echo "B carveout=default, api occupancy"; TSR_CARVEOUT=default timeout 200 python bench.py --only none --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
export TSR_OCCUPANCY=own
echo "C carveout=max, own occupancy b16"; timeout 200 python bench.py --only none --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
echo "D graph_step b64 own occupancy"; TOP=2 timeout 150 python tools/profile_step.py 64 2>&1 | tail -3 | cut -c1-200
