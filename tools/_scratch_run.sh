mkdir -p gpurun_out
export TSR_GRAPHS=0
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02g_ncu_launches_step.csv python tools/ncu_step.py 16 > gpurun_out/ncu_step_g.log 2>&1; tail -1 gpurun_out/ncu_step_g.log; wc -l gpurun_out/r02g_ncu_launches_step.csv
