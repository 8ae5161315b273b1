#!/bin/bash
F=multimodal-diagnosis-ham-spine_b200/csrc/norm.cu
cp $F /tmp/norm_orig.cu
for cfg in "0 0" "4 4" "5 5" "6 6"; do
  set -- $cfg
  cp /tmp/norm_orig.cu $F
  if [ "$1" != "0" ]; then
    sed -i -e "s/__launch_bounds__(256) bn_apply_kernel(/__launch_bounds__(256, $1) bn_apply_kernel(/" -e "s/__launch_bounds__(256) bn_bwd_apply_kernel(/__launch_bounds__(256, $2) bn_bwd_apply_kernel(/" $F
  fi
  python multimodal-diagnosis-ham-spine_b200/build.py > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-inference 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('occ=$cfg', d['ms_per_step'], d['value'])"
done
cp /tmp/norm_orig.cu $F
