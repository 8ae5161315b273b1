set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/bench_gemm_step.py l1_ > gpurun_out/r02f_gemm_step.jsonl 2>&1
timeout 600 python tools/bench_gemm_step.py ffn >> gpurun_out/r02f_gemm_step.jsonl 2>&1
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_boundary_gpu.py -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
timeout 600 python bench.py --no-gpu-baseline --no-cpu-baseline --no-inference > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
for cfg in "bf16 16" "bf16 0" "fp32 0"; do
  set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-inference --comm-dtype $1 --sm-reserve $2 > gpurun_out/r02f_bench_n2_$1_$2.json 2> gpurun_out/r02f_bench_n2_$1_$2.err
done
tail -3 gpurun_out/r02f_pytest.log; for f in gpurun_out/r02f_bench_n*.json; do echo $f; cut -c1-160 $f; done
