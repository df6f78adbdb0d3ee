timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "generator or golden or eval_mode or c1_fixture or large_image or pipelined or cli or kernel" 2>&1 | tail -2
TSR_CONV_VERBOSE=1 timeout 300 python tools/bench_infer.py 2> gpurun_out/v4.log | tail -1 | cut -c60-170; grep "out_mode=4" gpurun_out/v4.log | sort | uniq -c | cut -c1-220
timeout 300 python tools/profile_infer.py 2>&1 | tail -6 | cut -c1-90
