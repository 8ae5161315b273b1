set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_bn.py r02c > gpurun_out/r02c_bn.log 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_boundary_gpu.py -m gpu -q -x > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
timeout 600 python bench.py --no-gpu-baseline --no-cpu-baseline --no-inference > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
KEEP_REP="l1_conv3x3_fprop" bash tools/ncu_gemm.sh r02c l1_conv3x3_fprop l3_conv3x3_fprop l1_conv3x3_wgrad l1_dgrad_64_256 l1_1x1_64_256 out_fwd ffn1_fwd_deriv l3_conv3x3_dgrad_stat
ncu -i gpurun_out/r02c_gemm_l1_conv3x3_fprop.ncu-rep --page source --csv > gpurun_out/r02c_gemm_l1_conv3x3_fprop.source.csv 2>/dev/null
rm -f gpurun_out/r02c_gemm_l1_conv3x3_fprop.ncu-rep
tail -3 gpurun_out/r02c_pytest.log; cat gpurun_out/r02c_bench.json | cut -c1-200
