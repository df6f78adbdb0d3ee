mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "generator or golden or eval_mode or c1_fixture or large_image or pipelined or kernel or gan_step" 2>&1 | tail -3
TSR_CONV_VERBOSE=1 timeout 300 python tools/bench_infer.py 2> gpurun_out/verbose_infer2.log | tail -1; grep "out_mode=4" gpurun_out/verbose_infer2.log | sort | uniq -c
timeout 300 python tools/profile_infer.py 2>&1 | tail -6 | cut -c1-100
