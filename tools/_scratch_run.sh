mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "discriminator or vgg or kernel or gan_step or large_image" 2>&1 | tail -3
run() { timeout 400 python bench.py --only b64 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'b16', round(d['value']), d['ms_per_step'], 'b64', round(d['b64']['value']), d['b64']['ms_per_step'])"; }
run staged-first-layers
TSR_CONV_STAGED=0 run staged-off
