set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_boundary_gpu.py tests/test_connext.py -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
timeout 600 python tools/bench_gemm_step.py > gpurun_out/r02d_gemm_step.jsonl 2>&1
timeout 600 python bench.py --no-gpu-baseline --no-cpu-baseline --no-inference --dump-gemms > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden.py -m gpu -q > gpurun_out/r02d_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest2.log
tail -3 gpurun_out/r02d_pytest.log; tail -3 gpurun_out/r02d_pytest2.log; cat gpurun_out/r02d_bench.json | cut -c1-200
