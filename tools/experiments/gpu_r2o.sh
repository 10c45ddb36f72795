set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2o_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --config 2 --no-other-configs > gpurun_out/r2o_bench_c2.json 2> gpurun_out/r2o_bench_c2.err; echo "bench c2 rc=$?"; cut -c1-300 gpurun_out/r2o_bench_c2.json; tail -3 gpurun_out/r2o_bench_c2.err
