set -u
mkdir -p gpurun_out
for CR in 0.75 1.0 1.5; do for C in 1 8; do
  echo "== crit $CR, max bundle CTAs $C"
  BAYES_CRIT=$CR BAYES_MAX_BUNDLE_CTAS=$C python tools/schedule_report.py --config 4 --loci 5000 --top 3 > gpurun_out/sched_r3f_${CR}_${C}.txt 2>&1
  grep -E "wall clock|makespan|^  \(0|resident warps/SM over" gpurun_out/sched_r3f_${CR}_${C}.txt
  sed -n '/^longest/,/^first in launch/p' gpurun_out/sched_r3f_${CR}_${C}.txt | head -5
done; done
