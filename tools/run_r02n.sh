set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 4 5; do
timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-baseline --no-inference --steps 10 --timeline r02n_timeline_config$c.json > gpurun_out/r02n_bench_config$c.json 2> gpurun_out/r02n_bench_config$c.err
done
timeout 600 python -m pytest tests/test_oracle_vs_reference.py -q -x > gpurun_out/r02n_cpu_oracle_tests.log 2>&1
