#!/bin/bash
mkdir -p gpurun_out
export TSR_GRAPHS=0
timeout 200 python tools/ncu_step.py 16 > gpurun_out/ncu_plain.log 2>&1 &&
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01c.csv python tools/ncu_step.py 16 > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_plain.log; tail -1 gpurun_out/ncu_l.log; wc -l gpurun_out/launches_r01c.csv
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm_persistent -s 2 -c 3 \
    -f -o gpurun_out/prof_conv_persistent_r01c python tools/ncu_step.py 16 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 30 -c 2 \
    -f -o gpurun_out/prof_conv_trunk_r01c python tools/ncu_step.py 16 > gpurun_out/ncu_g.log 2>&1
tail -2 gpurun_out/ncu_g.log; ls -la gpurun_out/*r01c*
