#!/bin/bash
# One `ncu --set full` capture per hot non-GEMM kernel of the training step (eager step between cudaProfilerStart/Stop);
# keeps the raw-metric CSV only.  usage: tools/ncu_kernels_r02.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --profile-from-start off \
    --kernel-name "regex:$2" --launch-skip $3 --launch-count 1 -f -o $OUT/${TAG}_ncu_full_$1 python tools/profile_step.py > $OUT/${TAG}_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i $OUT/${TAG}_ncu_full_$1.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_$1.raw.csv 2>/dev/null
  rm -f $OUT/${TAG}_ncu_full_$1.ncu-rep $OUT/${TAG}_$1.log
}
# launch indices inside the step: layer-1 conv3 (256 channels, residual) = the largest BatchNorm tensors
cap bn_apply_l1 'bn_apply_kernel' 4
cap bn_bwd_apply_l1 'bn_bwd_apply_kernel' 49
cap bn_bwd_reduce_l1 'bn_bwd_reduce_kernel' 28
cap bn_bwd_apply_l3 'bn_bwd_apply_kernel' 20
cap ln_bwd_param 'ln_bwd_param_kernel' 5
cap attn_bwd 'attn_bwd_kernel' 3
