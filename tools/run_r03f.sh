set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for ds in 0 1; do
MDHS_DUAL_STREAM=$ds timeout 600 python bench.py --config 2 --no-cpu-baseline --no-gpu-baseline --no-inference --timeline r03f_timeline_ds$ds.json > gpurun_out/r03f_bench_config2_ds$ds.json 2> gpurun_out/r03f_bench_config2_ds$ds.err
python -c "
import json,sys;d=json.loads(open('gpurun_out/r03f_bench_config2_ds$ds.json').read().strip().splitlines()[-1]);print('ds$ds',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['gemm_ms_per_step'],d['roofline']['kernel_busy_ms_per_step'],d['final_loss'])"
done
MDHS_DUAL_STREAM=1 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_boundary_gpu.py -q -m gpu > gpurun_out/r03f_tests.log 2>&1
tail -5 gpurun_out/r03f_tests.log
