set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r03g_pytest_gpu.log 2>&1
tail -6 gpurun_out/r03g_pytest_gpu.log
for c in 2 3 4 5; do
timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/r03g_bench_config$c.json 2> gpurun_out/r03g_bench_config$c.err
python -c "
import json,sys;d=json.loads(open('gpurun_out/r03g_bench_config$c.json').read().strip().splitlines()[-1]);print($c,d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['gemm_ms_per_step'],d['final_loss'])"
done
