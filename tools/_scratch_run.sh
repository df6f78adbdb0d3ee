TSR_LIB_PATH=$PWD/torchsr_b200/lib/lib_prev.so timeout 120 python tools/microbench_epilogue.py 2>&1 | tail -1
timeout 120 python tools/microbench_epilogue.py 2>&1 | tail -1
TSR_LIB_PATH=$PWD/torchsr_b200/lib/lib_prev.so timeout 300 python tools/bench_infer.py 2>&1 | tail -1 | cut -c60-140
timeout 300 python tools/bench_infer.py 2>&1 | tail -1 | cut -c60-140
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "generator or golden or eval_mode or c1_fixture or large_image or kernel or vgg or discriminator" 2>&1 | tail -2
