mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 > gpurun_out/pytest_r2t.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2t.log | tail -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err; tail -c 300 gpurun_out/bench_r2t.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2t.json').read().strip().splitlines()[-1])
print('b16', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'b64', d['b64']['value'], 'infer', d['inference']['value'], d['inference']['e2e']['value'], d['inference']['e2e']['uint8_output']['value'], 'esrgan', d['esrgan']['value'], 'launches', d['launches_per_step'], 'eager x', d['gpu_eager_baseline']['speedup_over_best_variant'])
"
export TSR_GRAPHS=0
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02e_ncu_launches_infer.csv python tools/ncu_infer.py > gpurun_out/ncu_infer.log 2>&1; tail -1 gpurun_out/ncu_infer.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm_persistent -s 2 -c 3 \
    -f -o gpurun_out/r02e_full_infer_trunk python tools/ncu_infer.py > gpurun_out/ncu_f1.log 2>&1; tail -1 gpurun_out/ncu_f1.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm_persistent -s 34 -c 3 \
    -f -o gpurun_out/r02e_full_infer_tail python tools/ncu_infer.py > gpurun_out/ncu_f2.log 2>&1; tail -1 gpurun_out/ncu_f2.log
