set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2z_pytest.log; tail -6 gpurun_out/r2z_pytest.log
