#!/bin/bash
# ncu --set full of single GEMM cases of tools/bench_gemm_step.py (the eager warm-up launch #2 is captured).
# usage: tools/ncu_gemm.sh <tag> case1 case2 ...   -> gpurun_out/<tag>_gemm_<case>.raw.csv (+ .ncu-rep for KEEP_REP cases)
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for c in "$@"; do
  python tools/bench_gemm_step.py $c > /dev/null 2>&1 || { echo "$c failed"; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_tc_kernel --launch-skip 1 --launch-count 1 \
    -f -o $OUT/${TAG}_gemm_$c python tools/bench_gemm_step.py $c > $OUT/${TAG}_gemm_$c.log 2>&1
  echo "$c rc=$?"
  ncu -i $OUT/${TAG}_gemm_$c.ncu-rep --page raw --csv > $OUT/${TAG}_gemm_$c.raw.csv 2>/dev/null
  case " $KEEP_REP " in *" $c "*) ;; *) rm -f $OUT/${TAG}_gemm_$c.ncu-rep ;; esac
  rm -f $OUT/${TAG}_gemm_$c.log
done
