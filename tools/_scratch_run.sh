mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "generator or golden or eval_mode or c1_fixture or subpixel or residual_block or kernel" 2>&1 | tail -4
echo "--- default"; TSR_CONV_VERBOSE=1 timeout 300 python tools/bench_infer.py 2> gpurun_out/verbose_infer.log | tail -2; sort gpurun_out/verbose_infer.log | uniq -c | sort -rn | head -12
echo "--- HALO=0 (staged, im2col tiles)"; TSR_CONV_HALO=0 timeout 300 python tools/bench_infer.py 2>&1 | tail -1
echo "--- STAGED=0"; TSR_CONV_STAGED=0 timeout 300 python tools/bench_infer.py 2>&1 | tail -1
timeout 300 python tools/profile_infer.py > gpurun_out/profile_infer.log 2>&1; tail -48 gpurun_out/profile_infer.log | cut -c1-110
