set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -k "im2col or tta or stem or conv" > gpurun_out/r03b_tests.log 2>&1
tail -4 gpurun_out/r03b_tests.log
timeout 600 python bench.py --config 2 --no-cpu-baseline --no-gpu-baseline --timeline r03b_timeline_config2.json > gpurun_out/r03b_bench_config2.json 2> gpurun_out/r03b_bench_config2.err
python tools/timeline_top.py gpurun_out/r03b_timeline_config2.json 30
