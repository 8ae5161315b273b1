set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
run() {  # tag, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-inference $2 > gpurun_out/r02l_bench_n${N}_$1.json 2> gpurun_out/r02l_bench_n${N}_$1.err
}
NCCL_PROTO=Simple run bwd_simple "--overlap-comm backward --timeline r02l_timeline_n${N}_bwd_simple.json"
NCCL_PROTO=Simple NCCL_ALGO=NVLS run bwd_nvls "--overlap-comm backward"
NCCL_PROTO=Simple run none_simple "--overlap-comm none"
NCCL_PROTO=Simple run bwd_simple_b0 "--overlap-comm backward --bert-bucket-layers 0"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING run bwd_info "--overlap-comm backward --steps 3 --warmup 3"
grep -i "nvls\|algo\|proto\|channel" gpurun_out/r02l_bench_n${N}_bwd_info.err | head -40 > gpurun_out/r02l_nccl_info.txt
for f in gpurun_out/r02l_bench_n${N}_*.json; do echo $f; cut -c1-140 $f; done
