TRACE=0 timeout 300 python tools/trace_fused.py 2>&1 | tail -2
timeout 300 python tools/trace_fused.py 2>&1 | tail -12
timeout 200 python tools/bench_programs.py 16 2>&1 | tail -5 | cut -c1-100
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "stage or oracle or golden or pair or residual" 2>&1 | tail -3
