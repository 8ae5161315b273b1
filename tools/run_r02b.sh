set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_fullsize_gpu.py > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s > gpurun_out/r02b_pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_full.log
timeout 600 python bench.py --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
timeout 300 python tools/bench_bn.py r02b > gpurun_out/r02b_bn.log 2>&1
timeout 600 python tools/bench_gemm_step.py > gpurun_out/r02b_gemm_step.jsonl 2>&1
tail -3 gpurun_out/r02b_pytest.log; tail -3 gpurun_out/r02b_pytest_full.log; cat gpurun_out/r02b_bench.json | cut -c1-300
