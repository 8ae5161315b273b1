# Final-build measurements of round 2 (after the depthwise / attention / LayerNorm / im2col changes).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
for c in 1 3 4 5; do
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline --timeline r02_timeline_config$c.json > gpurun_out/r02_bench_config$c.json 2> gpurun_out/r02_bench_config$c.err
done
timeout 600 python bench.py --config 2 --seq 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_config2_s128.json 2> gpurun_out/r02_bench_config2_s128.err
timeout 300 python tools/one_attention.py > gpurun_out/r02_attention_microbench.log 2>&1
timeout 300 python tools/one_dwconv.py > gpurun_out/r02_dwconv_microbench.log 2>&1
# ncu launch list of one config-2 step with DRAM bytes (after the same command ran clean)
timeout 300 python tools/profile_step.py > gpurun_out/r02_profile_step_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step_b128.csv python tools/profile_step.py > gpurun_out/r02_ncu_launches.log 2>&1
# ncu --set full of the new kernels (each target ran clean just above)
for k in attn_fwd_long attn_bwd_dq attn_bwd_dkv; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 2 --launch-count 1 -f -o gpurun_out/tmp_$k python tools/one_attention.py 256 > gpurun_out/r02_ncu_$k.log 2>&1
ncu -i gpurun_out/tmp_$k.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_$k.raw.csv 2>/dev/null
done
for k in dwconv7_tma dwconv7_wgrad_tma; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 3 --launch-count 1 -f -o gpurun_out/tmp_$k python tools/one_dwconv.py 1 > gpurun_out/r02_ncu_$k.log 2>&1
ncu -i gpurun_out/tmp_$k.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_$k.raw.csv 2>/dev/null
done
rm -f gpurun_out/tmp_*.ncu-rep
for f in gpurun_out/r02_bench_*.json; do echo $f; cut -c1-160 $f; done
cat gpurun_out/r02_attention_microbench.log gpurun_out/r02_dwconv_microbench.log
