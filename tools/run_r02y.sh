set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02y_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02y_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err
tail -c 600 gpurun_out/r02y_bench_n1.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; tail -2 gpurun_out/r02y_smoke.log
