mkdir -p gpurun_out
for f in 1 0; do echo "== TSR_BN_FUSE=$f"; TSR_BN_FUSE=$f timeout 200 python tools/bench_programs.py 16 2>&1 | tail -8 | cut -c1-200; done
TIMELINE=gpurun_out/timeline_r2b.csv TOP=25 timeout 200 python tools/profile_step.py 16 2>&1 | tail -30 | cut -c1-200
