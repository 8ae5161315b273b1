set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
run() {  # tag, extra args, env
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-inference $2 > gpurun_out/r02k_bench_n${N}_$1.json 2> gpurun_out/r02k_bench_n${N}_$1.err
}
run pipeline "--overlap-comm pipeline --timeline r02k_timeline_n${N}_pipeline.json"
run backward "--overlap-comm backward --timeline r02k_timeline_n${N}_backward.json"
run none "--overlap-comm none"
MDHS_COMM_SMS=16 NCCL_MAX_CTAS=16 run pipeline_excl "--overlap-comm pipeline"
run backward_fp32 "--overlap-comm backward --comm-dtype fp32 --bert-bucket-layers 0"
for f in gpurun_out/r02k_bench_n${N}_*.json; do echo $f; cut -c1-140 $f; done
