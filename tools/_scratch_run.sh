for i in 1 2; do
echo "prev"; TSR_LIB_PATH=$PWD/torchsr_b200/lib/lib_prev.so timeout 300 python tools/bench_infer.py 2>&1 | tail -1 | cut -c60-140
echo "new";  timeout 300 python tools/bench_infer.py 2>&1 | tail -1 | cut -c60-140
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
