set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_fullsize_gpu.py > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s > gpurun_out/r02a_pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_full.log
timeout 600 python bench.py --dump-gemms > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
timeout 300 python tools/bench_bn.py r02a_8_2 > gpurun_out/r02a_bn.log 2>&1
MDHS_BN_REDUCE=8,4 timeout 300 python tools/bench_bn.py r02a_8_4 >> gpurun_out/r02a_bn.log 2>&1
MDHS_BN_REDUCE=16,1 timeout 300 python tools/bench_bn.py r02a_16_1 >> gpurun_out/r02a_bn.log 2>&1
MDHS_BN_REDUCE=16,2 timeout 300 python tools/bench_bn.py r02a_16_2 >> gpurun_out/r02a_bn.log 2>&1
tail -3 gpurun_out/r02a_pytest.log; tail -3 gpurun_out/r02a_pytest_full.log; cat gpurun_out/r02a_bench.json | cut -c1-600
