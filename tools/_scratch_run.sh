for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q --timeout=600 -x -k "c1_fixture or large_image or inference_plan or eval_mode" 2>&1 | tail -1; done
timeout 300 python tools/bench_infer.py 2>&1 | tail -1 | cut -c60-170
