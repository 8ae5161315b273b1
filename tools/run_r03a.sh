set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_connext.py tests/test_model_gpu.py -q -m gpu > gpurun_out/r03a_tests.log 2>&1
tail -8 gpurun_out/r03a_tests.log
for c in 2 4; do
timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline --timeline r03a_timeline_config$c.json > gpurun_out/r03a_bench_config$c.json 2> gpurun_out/r03a_bench_config$c.err
python -c "
import json,sys;d=json.loads(open('gpurun_out/r03a_bench_config$c.json').read().strip().splitlines()[-1]);print($c,d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'])"
done
