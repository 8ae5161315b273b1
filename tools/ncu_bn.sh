#!/bin/bash
OUT=gpurun_out
cap() {
  timeout 300 ncu --set full --clock-control none --profile-from-start off --kernel-name "regex:$2" --launch-skip $3 --launch-count 1 -f -o $OUT/r01y_$1 python tools/profile_step.py > /dev/null 2>&1
  ncu -i $OUT/r01y_$1.ncu-rep --page raw --csv > $OUT/r01y_$1.raw.csv 2>/dev/null; rm -f $OUT/r01y_$1.ncu-rep
}
cap bn_apply 'bn_apply_kernel' 4
cap bn_bwd_apply 'bn_bwd_apply_kernel' 46
cap bn_bwd_reduce 'bn_bwd_reduce_kernel' 46
cap bn_bwd_reduce_l1 'bn_bwd_reduce_kernel' 50
