set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
BAYES_B200_LIB=$PWD/bayesembler_b200/libbayes_phases.so python tools/tile_phases.py 4 3000 2>&1 | grep -v "^tile phases" | tee gpurun_out/r3m_tile_phases.txt
echo "== 5000 loci"; python tools/schedule_report.py --config 4 --loci 5000 --top 3 2>&1 | grep -E "wall clock|makespan|resident warps/SM over"
echo "== 2500 loci"; python tools/schedule_report.py --config 4 --loci 2500 --top 3 2>&1 | grep -E "wall clock|makespan|resident warps/SM over"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/r3m_bench.json 2> gpurun_out/r3m_bench.err; python -c "
import json; l=json.load(open('gpurun_out/r3m_bench.json')); print('c4 N=1: %.1f ms/step, e2e %.1f ms/step, launches %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['gpu_launches']))"
python tools/schedule_report.py --config 2 --top 3 2>&1 | grep -E "wall clock|makespan"
