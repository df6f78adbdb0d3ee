timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "generator or golden or eval_mode or c1_fixture or large_image or pipelined or cli or inference_plan or packed_weights" 2>&1 | tail -3
timeout 300 python tools/bench_infer.py 2>&1 | tail -1 | cut -c60-170
