set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_connext.py -q -m gpu > gpurun_out/r02s_test_connext.log 2>&1
tail -5 gpurun_out/r02s_test_connext.log
timeout 300 python tools/one_dwconv.py > gpurun_out/r02s_dwconv.log 2>&1
cat gpurun_out/r02s_dwconv.log
timeout 600 python bench.py --config 4 --steps 20 --warmup 5 > gpurun_out/r02s_bench_config4.json 2> gpurun_out/r02s_bench_config4.err
tail -c 1500 gpurun_out/r02s_bench_config4.json
