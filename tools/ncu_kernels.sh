#!/bin/bash
# One `ncu --set full` capture per hot kernel of the training step (eager step between cudaProfilerStart/Stop).
# usage: tools/ncu_kernels.sh <tag>      -> gpurun_out/<tag>_<kernel>.ncu-rep
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python tools/profile_step.py > /dev/null 2>&1 || { echo "step failed"; exit 1; }
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    --kernel-name "regex:$2" --launch-skip $3 --launch-count 1 -f -o $OUT/${TAG}_$1 python tools/profile_step.py > $OUT/${TAG}_$1.log 2>&1
  echo "$1 rc=$?"
  # the reports are big (source import): keep the raw-metric CSV, drop the .ncu-rep unless KEEP_REP names this kernel
  ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1.raw.csv 2>/dev/null
  case " $KEEP_REP " in *" $1 "*) ;; *) rm -f $OUT/${TAG}_$1.ncu-rep ;; esac
  rm -f $OUT/${TAG}_$1.log
}
cap ln_bwd 'ln_bwd_kernel' 5
cap ln_fwd 'ln_fwd_kernel' 5
cap attn_bwd 'attn_bwd_kernel' 3
cap attn_fwd 'attn_fwd_kernel' 3
cap bn_apply 'bn_apply_kernel' 4
cap bn_bwd_apply 'bn_bwd_apply_kernel' 46
cap bn_bwd_reduce 'bn_bwd_reduce_kernel' 46
cap col_stats 'col_stats_kernel' 4
cap im2col_stem 'im2col_nchw_f32_kernel' 0
cap adam 'adam_flat_kernel' 0
cap gemm_ffn1 'gemm_tc_kernel' 8
ls -la $OUT/${TAG}_*
