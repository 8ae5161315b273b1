set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_boundary_gpu.py tests/test_golden.py -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
timeout 600 python bench.py --no-gpu-baseline --no-cpu-baseline --no-inference --timeline r02j_timeline_n1.json > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-inference --timeline r02j_timeline_n2.json > gpurun_out/r02j_bench_n2.json 2> gpurun_out/r02j_bench_n2.err
tail -3 gpurun_out/r02j_pytest.log; for f in gpurun_out/r02j_bench_n*.json; do echo $f; cut -c1-160 $f; done
