mkdir -p gpurun_out
export TSR_GRAPHS=0
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02c_ncu_launches_infer.csv python tools/ncu_infer.py > gpurun_out/ncu_infer.log 2>&1; tail -1 gpurun_out/ncu_infer.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02c_ncu_launches_step_b64.csv python tools/ncu_step.py 64 > gpurun_out/ncu_b64.log 2>&1; tail -1 gpurun_out/ncu_b64.log
unset TSR_GRAPHS
TIMELINE=gpurun_out/timeline_b64_r2n.csv TOP=14 timeout 300 python tools/profile_step.py 64 2>&1 | tail -18 | cut -c1-170
