# Final-build bench lines of round 2 (dual-stream encoders + dynamic GEMM scheduling on).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --timeline r02_timeline_config2.json > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
for c in 1 3 4 5; do
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline --timeline r02_timeline_config$c.json > gpurun_out/r02_bench_config$c.json 2> gpurun_out/r02_bench_config$c.err
done
timeout 600 python bench.py --config 2 --seq 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_config2_s128.json 2> gpurun_out/r02_bench_config2_s128.err
MDHS_DUAL_STREAM=0 MDHS_GEMM_DYNAMIC=0 timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/r02_bench_n1_single_stream_static.json 2> /dev/null
for f in gpurun_out/r02_bench_*.json; do python -c "
import json,sys;d=json.loads(open('$f').read().strip().splitlines()[-1]);r=d['roofline'];print('$f'.split('/')[-1],d['value'],d['ms_per_step'],d['e2e']['value'],r['achieved'],r['frac'],r.get('single_stream_ms_per_step'),d['clocks']['sm_mhz'])"; done
