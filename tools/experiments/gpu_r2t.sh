set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2t_parity.log; tail -5 gpurun_out/r2t_parity.log
timeout 600 python tools/rows_timing.py --loci 3 --iters 2000 --together > gpurun_out/r2t_rows_timing.txt 2>&1; cat gpurun_out/r2t_rows_timing.txt
