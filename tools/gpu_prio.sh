#!/bin/bash
for cfg in "-1 0 -1 x" "-1 -1 -1 x" "0 0 0 x" "-1 -2 -1 x" "-2 -1 -2 x" "-1 -1 -1 0" "-1 0 -1 0"; do
  set -- $cfg
  W=""; [ "$4" != "x" ] && W="TSR_WGRAD_PRIO=$4"
  echo "SIDE=$1 VGG=$2 MAIN=$3 WGRAD=$4: $(env TSR_PRIO_SIDE=$1 TSR_PRIO_VGG=$2 TSR_PRIO_MAIN=$3 $W timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-160)"
done
