set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/one_attention.py 256 > gpurun_out/r02u_attn.log 2>&1 || exit 1
for k in attn_fwd_long attn_bwd_dq attn_bwd_dkv; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02u_$k python tools/one_attention.py 256 > gpurun_out/r02u_ncu_$k.log 2>&1
ncu -i gpurun_out/r02u_$k.ncu-rep --page raw --csv > gpurun_out/r02u_$k.raw.csv 2>/dev/null
ncu -i gpurun_out/r02u_$k.ncu-rep --page source --csv > gpurun_out/r02u_$k.source.csv 2>/dev/null
done
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/r02u_attn.log
