run() { timeout 400 python bench.py --only b64 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'b16', round(d['value']), d['ms_per_step'], 'b64', round(d['b64']['value']), d['b64']['ms_per_step'])"; }
run default
TSR_RING_KB=190 run ring190
TSR_PERSIST_MIN_TILES_X10=12 run persist1.2
TSR_RING_KB=190 TSR_PERSIST_MIN_TILES_X10=12 run both
