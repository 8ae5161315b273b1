set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# 1. the driver's command line, with every baseline
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
# 2. reference arms
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference_cpu.json 2> gpurun_out/r02_bench_reference_cpu.err
# 3. the other BASELINE configurations (device-timed, no baselines)
for c in 1 3 4 5; do
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_config$c.json 2> gpurun_out/r02_bench_config$c.err
done
timeout 600 python bench.py --config 2 --seq 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_config2_s128.json 2> gpurun_out/r02_bench_config2_s128.err
timeout 900 python bench.py --impl torch_gpu --config 5 --steps 5 --warmup 3 > gpurun_out/r02_bench_torch_gpu_config5.json 2> gpurun_out/r02_bench_torch_gpu_config5.err
# 4. BN / GEMM micro-benchmarks (final kernels)
timeout 300 python tools/bench_bn.py r02_final > gpurun_out/r02_bn_final.log 2>&1
timeout 600 python tools/bench_gemm_step.py > gpurun_out/r02_gemm_microbench_final.jsonl 2>&1
# 5. ncu launch list of one step with DRAM bytes (after the same command ran clean)
timeout 300 python tools/profile_step.py > gpurun_out/r02_profile_step_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step_b128.csv python tools/profile_step.py > gpurun_out/r02_ncu_launches.log 2>&1
# 6. ncu --set full of the BatchNorm kernels of the step
KEEP_REP="" bash tools/ncu_kernels_r02.sh r02 > gpurun_out/r02_ncu_kernels.log 2>&1
for f in gpurun_out/r02_bench_*.json; do echo $f; cut -c1-160 $f; done
