set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for ds in 0 1; do
for c in 2 5; do
MDHS_DUAL_STREAM=$ds timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/r03d_bench_config${c}_ds$ds.json 2> gpurun_out/r03d_bench_config${c}_ds$ds.err
python -c "
import json,sys;d=json.loads(open('gpurun_out/r03d_bench_config${c}_ds$ds.json').read().strip().splitlines()[-1]);print('ds$ds',$c,d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['gemm_ms_per_step'],d['clocks'],d['final_loss'])"
done
done
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_boundary_gpu.py tests/test_fullsize_gpu.py tests/test_golden.py tests/test_connext.py -q -m gpu > gpurun_out/r03d_tests.log 2>&1
tail -5 gpurun_out/r03d_tests.log
