set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" > gpurun_out/r02x_test_attn.log 2>&1
tail -15 gpurun_out/r02x_test_attn.log
timeout 300 python tools/one_attention.py > gpurun_out/r02x_attn.log 2>&1
cat gpurun_out/r02x_attn.log
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r02x_bench_config5.json 2> gpurun_out/r02x_bench_config5.err
tail -c 1200 gpurun_out/r02x_bench_config5.json
