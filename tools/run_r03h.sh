set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r03h_bench_n1.json 2> gpurun_out/r03h_bench_n1.err
tail -c 2500 gpurun_out/r03h_bench_n1.json; tail -3 gpurun_out/r03h_bench_n1.err
