set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --no-gpu-baseline --no-cpu-baseline --no-inference --timeline r02g_timeline_n1.json > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-inference --timeline r02g_timeline_n2.json > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err
timeout 600 python tools/bench_gemm_step.py conv3x3 > gpurun_out/r02g_gemm_step.jsonl 2>&1
for f in gpurun_out/r02g_bench_n*.json; do echo $f; cut -c1-160 $f; done
