set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/one_dwconv.py > gpurun_out/r02o_dwconv.log 2>&1
timeout 300 python tools/one_dwconv.py 1 > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:dwconv7 --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02o_dwconv_fwd python tools/one_dwconv.py 1 > gpurun_out/r02o_ncu1.log 2>&1
ncu -i gpurun_out/r02o_dwconv_fwd.ncu-rep --page raw --csv > gpurun_out/r02o_dwconv_fwd.raw.csv 2>/dev/null
ncu -i gpurun_out/r02o_dwconv_fwd.ncu-rep --page source --csv > gpurun_out/r02o_dwconv_fwd.source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:dwconv7_wgrad --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02o_dwconv_wgrad python tools/one_dwconv.py 1 > gpurun_out/r02o_ncu2.log 2>&1
ncu -i gpurun_out/r02o_dwconv_wgrad.ncu-rep --page raw --csv > gpurun_out/r02o_dwconv_wgrad.raw.csv 2>/dev/null
ncu -i gpurun_out/r02o_dwconv_wgrad.ncu-rep --page source --csv > gpurun_out/r02o_dwconv_wgrad.source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/r02o_dwconv.log
