for v in 10 20 40 100000; do
  echo "PERSIST_MIN_X10=$v: $(TSR_PERSIST_MIN_TILES_X10=$v timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-175)"
done
for v in 10 40; do
  echo "infer PERSIST_MIN_X10=$v: $(TSR_PERSIST_MIN_TILES_X10=$v timeout 100 python tools/bench_infer.py 1 512 2>&1 | tail -1 | cut -c60-140)"
done
