set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/r03i_bench_n2_dual_dyn.json 2> gpurun_out/r03i_bench_n2_dual_dyn.err
MDHS_DUAL_STREAM=0 MDHS_GEMM_DYNAMIC=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/r03i_bench_n2_single_static.json 2> gpurun_out/r03i_bench_n2_single_static.err
MDHS_DUAL_STREAM=0 MDHS_GEMM_DYNAMIC=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/r03i_bench_n2_single_dyn.json 2> gpurun_out/r03i_bench_n2_single_dyn.err
for f in dual_dyn single_static single_dyn; do python -c "
import json,sys;d=json.loads(open('gpurun_out/r03i_bench_n2_$f.json').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['e2e']['value'],d['final_loss'])"; tail -2 gpurun_out/r03i_bench_n2_$f.err; done
