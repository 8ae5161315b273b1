set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py tests/test_golden.py tests/test_kernels_gpu.py -q -m gpu > gpurun_out/r03c_tests.log 2>&1
tail -5 gpurun_out/r03c_tests.log
for c in 2 5; do
timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline > gpurun_out/r03c_bench_config$c.json 2> gpurun_out/r03c_bench_config$c.err
python -c "
import json,sys;d=json.loads(open('gpurun_out/r03c_bench_config$c.json').read().strip().splitlines()[-1]);print($c,d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['gemm_ms_per_step'])"
done
