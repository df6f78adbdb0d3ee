#!/bin/bash
# Round-2 ncu evidence: (1) plain run, (2) launch list of ONE training step (gpu__time_duration per kernel),
# (3) --set full captures of the kernel families the roofline / hbm_kernels numbers are about.
TAG=${1:-r02}
mkdir -p gpurun_out
export TSR_GRAPHS=0
timeout 300 python tools/ncu_step.py 16 > gpurun_out/ncu_plain_$TAG.log 2>&1 || { tail -5 gpurun_out/ncu_plain_$TAG.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_ncu_launches_step.csv python tools/ncu_step.py 16 > gpurun_out/ncu_l_$TAG.log 2>&1
tail -1 gpurun_out/ncu_l_$TAG.log; wc -l gpurun_out/${TAG}_ncu_launches_step.csv
full() {  # name regex skip count
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 \
      -f -o gpurun_out/${TAG}_full_$1 python tools/ncu_step.py 16 > gpurun_out/ncu_f_$1_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_f_$1_$TAG.log
}
full adam adam_pack_kernel 0 3
full bn_bwd_apply bn_bwd_apply_kernel 10 3
full bn_act bn_act_kernel 0 3
full wgrad conv_wgrad_kernel 4 4
full conv_persistent conv_igemm_persistent 2 4
full conv_trunk "conv_igemm_kernel" 20 3
ls -la gpurun_out/${TAG}_full_* | head
