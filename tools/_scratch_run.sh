timeout 900 python -m pytest tests -m gpu -q --timeout=600 -s -k "large_image" 2>&1 | grep -E "large-image|passed|failed"
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "generator or golden or eval_mode or c1_fixture or pipelined or cli or kernel" 2>&1 | tail -2
