set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3b_parity.log; tail -4 gpurun_out/r3b_parity.log
BAYES_B200_LIB=$PWD/bayesembler_b200/libbayes_phases.so python tools/tile_phases.py 4 3000 2>&1 | grep -v "^tile phases" | tee gpurun_out/r3b_tile_phases.txt
BAYES_B200_LIB=$PWD/bayesembler_b200/libbayes_phases.so BAYES_FORCE_WARPS=256 python tools/run_one.py --config 4 --locus 29910 --iters 3000 | grep "tile phases" | head -3
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; python -c "
import json; l=json.load(open('gpurun_out/r3b_bench.json')); print('c4 N=1: %.1f ms/step, e2e %.1f ms/step, launches %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['gpu_launches']))"
