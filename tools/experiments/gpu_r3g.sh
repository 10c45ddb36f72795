set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster_bundle or batch_mixed" 2>&1 | tail -3
echo "== 5000 loci"; python tools/schedule_report.py --config 4 --loci 5000 --top 3 > gpurun_out/sched_r3g_5000.txt 2>&1
grep -E "wall clock|makespan|^  \(0|resident warps/SM over" gpurun_out/sched_r3g_5000.txt; sed -n '/^longest/,/^first in launch/p' gpurun_out/sched_r3g_5000.txt | head -5
echo "== 10000 loci"; python tools/schedule_report.py --config 4 --loci 10000 --top 3 > gpurun_out/sched_r3g_10000.txt 2>&1
grep -E "wall clock|makespan|resident warps/SM over" gpurun_out/sched_r3g_10000.txt
echo "== 10000 loci, crit 0.5, no bundle clusters"; BAYES_CRIT=0.5 BAYES_MAX_BUNDLE_CTAS=1 python tools/schedule_report.py --config 4 --loci 10000 --top 3 > gpurun_out/sched_r3g_10000_old.txt 2>&1
grep -E "wall clock|makespan|resident warps/SM over" gpurun_out/sched_r3g_10000_old.txt
for CR in 1.0 0.5; do BAYES_CRIT=$CR python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/r3g_bench_$CR.json 2> gpurun_out/r3g_bench_$CR.err; python -c "
import json; l=json.load(open('gpurun_out/r3g_bench_$CR.json')); print('c4 N=1 crit $CR: %.1f ms/step, e2e %.1f ms/step, launches %d' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['gpu_launches']))"; done
