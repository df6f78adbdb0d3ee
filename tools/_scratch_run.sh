for b in 1 2 4; do timeout 200 python tools/bench_infer.py $b 512 2>&1 | tail -1; done
timeout 200 python tools/bench_infer.py 1 1024 2>&1 | tail -1
timeout 200 python tools/bench_infer.py 1 256 2>&1 | tail -1
