#!/bin/bash
mkdir -p gpurun_out
export TSR_GRAPHS=0
timeout 200 python tools/ncu_step.py 16 > gpurun_out/ncu_plain.log 2>&1 &&
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01b.csv python tools/ncu_step.py 16 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_plain.log; tail -2 gpurun_out/ncu_l.log; wc -l gpurun_out/launches_r01b.csv
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm -s 60 -c 3 \
    -f -o gpurun_out/prof_conv_r01b python tools/ncu_step.py 16 > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log; ls -la gpurun_out/*.ncu-rep
