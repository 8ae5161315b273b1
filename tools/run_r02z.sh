set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 1 3 4 5; do
timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline --timeline r02z_timeline_config$c.json > gpurun_out/r02z_bench_config$c.json 2> gpurun_out/r02z_bench_config$c.err
done
timeout 600 python bench.py --config 2 --seq 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02z_bench_config2_s128.json 2> gpurun_out/r02z_bench_config2_s128.err
for f in gpurun_out/r02z_bench_*.json; do python -c "
import json,sys;d=json.loads(open('$f').read().strip().splitlines()[-1]);print('$f',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'])"; done
