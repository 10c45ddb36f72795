set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast or full_chain or batch or cluster" 2>&1 | tail -2
BAYES_B200_LIB=$PWD/bayesembler_b200/libbayes_phases.so python tools/tile_phases.py 4 3000 2>&1 | grep -v "^tile phases" | tee gpurun_out/r3n_tile_phases.txt
BAYES_B200_LIB=$PWD/bayesembler_b200/libbayes_phases.so BAYES_FORCE_WARPS=256 python tools/run_one.py --config 4 --locus 10143 --iters 3000 | grep "tile phases" | head -2
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err; python -c "
import json; l=json.load(open('gpurun_out/r3n_bench.json')); print('c4 N=1: %.1f ms/step, e2e %.1f ms/step, launches %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['gpu_launches']))"
