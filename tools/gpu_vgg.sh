#!/bin/bash
timeout 200 python -m pytest tests/test_modules_gpu.py -m gpu -x -q -s -k "vgg or gan_step" 2>&1 | grep -E "vgg|passed|failed|Error|error" | tail -8
for impl in torch b200; do
  echo "VGG impl $impl: $(TORCHSR_VGG_IMPL=$impl timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-170)"
done
