set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_fullsize_gpu.py > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s > gpurun_out/r02m_pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest_full.log
timeout 600 python bench.py --no-gpu-baseline --no-cpu-baseline --timeline r02m_timeline_n1.json > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02m_smoke.log
tail -3 gpurun_out/r02m_pytest.log; tail -3 gpurun_out/r02m_pytest_full.log; tail -2 gpurun_out/r02m_smoke.log; cut -c1-200 gpurun_out/r02m_bench_n1.json
