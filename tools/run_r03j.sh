set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/r03j_bench_n8.json 2> gpurun_out/r03j_bench_n8.err
python -c "
import json,sys;d=json.loads(open('gpurun_out/r03j_bench_n8.json').read().strip().splitlines()[-1]);print('n8',d['value'],d['ms_per_step'],d['e2e']['value'],d['final_loss'])"; tail -2 gpurun_out/r03j_bench_n8.err
