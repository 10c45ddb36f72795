set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster_bundle or batch_mixed or full_chain" 2>&1 | tail -8 > gpurun_out/r3a_parity.log; tail -5 gpurun_out/r3a_parity.log
BAYES_B200_LIB=$PWD/bayesembler_b200/libbayes_phases.so python tools/tile_phases.py 4 3000 2>&1 | grep -v "^tile phases.*256 threads" | awk '!seen[$0]++' | tee gpurun_out/r3a_tile_phases.txt
