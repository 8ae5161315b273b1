set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -m gpu -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
timeout 600 python tools/bench_gemm_step.py > gpurun_out/r02e_gemm_step.jsonl 2>&1
timeout 600 python bench.py --no-gpu-baseline --no-cpu-baseline --no-inference --dump-gemms > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
KEEP_REP="l1_conv3x3_fprop" bash tools/ncu_gemm.sh r02e l1_conv3x3_fprop l1_1x1_64_256 l1_dgrad_64_256
ncu -i gpurun_out/r02e_gemm_l1_conv3x3_fprop.ncu-rep --page source --csv > gpurun_out/r02e_gemm_l1_conv3x3_fprop.source.csv 2>/dev/null
rm -f gpurun_out/r02e_gemm_l1_conv3x3_fprop.ncu-rep
tail -3 gpurun_out/r02e_pytest.log; cat gpurun_out/r02e_bench.json | cut -c1-200
