set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err; python -c "
import json; l=json.load(open('gpurun_out/r3c_bench.json')); print('c4 N=1: %.1f ms/step, e2e %.1f ms/step, launches %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['gpu_launches']))"
BAYES_MAX_BUNDLE_CTAS=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/r3c_bench_noclu.json 2> gpurun_out/r3c_bench_noclu.err; python -c "
import json; l=json.load(open('gpurun_out/r3c_bench_noclu.json')); print('c4 N=1 without bundle clusters: %.1f ms/step, e2e %.1f ms/step, launches %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['gpu_launches']))"
python tools/schedule_report.py --top 3 > gpurun_out/sched_r3c_c4.txt 2>&1; head -12 gpurun_out/sched_r3c_c4.txt; grep "resident warps/SM over" gpurun_out/sched_r3c_c4.txt
