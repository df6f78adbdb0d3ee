mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --only b64 > gpurun_out/scale8_r2t.json 2> gpurun_out/scale8_r2t.err; tail -c 300 gpurun_out/scale8_r2t.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale8_r2t.json') if l.startswith('{')][-1])
print('n8 b16', d['value'], d['ms_per_step'], 'dp_parity', d.get('dp_parity'), 'b64', d.get('b64',{}).get('value'))
"
