#!/bin/bash
# Round-end verification on one B200: GPU test-suite, smoke, the bench lines of every configuration, ncu launch list
# of one step and one --set full capture of the conv kernels. Outputs under gpurun_out/ (tag = $1).
TAG=${1:-r01d}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --steps 50 --warmup 10 2>&1 | tail -1 > gpurun_out/bench_$TAG.json; cut -c1-330 gpurun_out/bench_$TAG.json
echo "B64: $(timeout 120 python bench.py --batch 64 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-170)" | tee gpurun_out/bench_b64_$TAG.log
timeout 100 python tools/bench_infer.py 1 512 2>&1 | tail -1 | tee gpurun_out/bench_infer_$TAG.log
timeout 200 python tools/bench_esrgan.py 2>&1 | tail -1 | tee gpurun_out/bench_esrgan_$TAG.log
export TSR_GRAPHS=0
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$TAG.csv python tools/ncu_step.py 16 > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log; wc -l gpurun_out/launches_$TAG.csv
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm -s 60 -c 4 \
    -f -o gpurun_out/prof_conv_$TAG python tools/ncu_step.py 16 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log; ls -la gpurun_out/*$TAG*
