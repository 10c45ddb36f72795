set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster or fast or full_chain or batch" 2>&1 | tail -15 > gpurun_out/r2x_parity.log; tail -5 gpurun_out/r2x_parity.log
for L in 11 459; do for N in 1 4 16; do BAYES_B200_LIB=$PWD/bayesembler_b200/libbayes_phases.so BAYES_FORCE_CTAS=$N python tools/run_one.py --config 3 --locus $L --iters 3000 | grep "rows phases"; done; done
timeout 600 python tools/rows_timing.py --loci 3 --iters 3000 --together > gpurun_out/r2x_rows_timing.txt 2>&1; cat gpurun_out/r2x_rows_timing.txt
