set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/r03e_test_gemm.log 2>&1
tail -6 gpurun_out/r03e_test_gemm.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv or gemm or batchnorm" > gpurun_out/r03e_test_k.log 2>&1
tail -3 gpurun_out/r03e_test_k.log
for dy in 0 1; do
MDHS_GEMM_DYNAMIC=$dy timeout 600 python bench.py --config 2 --no-cpu-baseline --no-gpu-baseline --no-inference > gpurun_out/r03e_bench_config2_dyn$dy.json 2> gpurun_out/r03e_bench_config2_dyn$dy.err
python -c "
import json,sys;d=json.loads(open('gpurun_out/r03e_bench_config2_dyn$dy.json').read().strip().splitlines()[-1]);print('dyn$dy',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['gemm_ms_per_step'],d['final_loss'])"
done
