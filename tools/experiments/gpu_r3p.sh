set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "batch_mixed or full_size or concurrent" 2>&1 | tail -2
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/r3p_bench.json 2> gpurun_out/r3p_bench.err; python -c "
import json; l=json.load(open('gpurun_out/r3p_bench.json')); print('c4 N=1: %.1f ms/step, e2e %.1f ms/step' % (l['ms_per_step'], l['e2e']['ms_per_step'])); print(l['e2e']['host_phases_s_rank0'])"
