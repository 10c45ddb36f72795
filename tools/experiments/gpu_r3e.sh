set -u
mkdir -p gpurun_out
for C in 1 2 4 8; do
  echo "== max bundle CTAs $C"
  BAYES_MAX_BUNDLE_CTAS=$C python tools/schedule_report.py --config 4 --loci 5000 --top 4 > gpurun_out/sched_r3e_c$C.txt 2>&1
  grep -E "wall clock|makespan|launches per run|^  \(0|resident warps/SM over" gpurun_out/sched_r3e_c$C.txt
  sed -n '/^longest/,/^first in launch/p' gpurun_out/sched_r3e_c$C.txt | head -7
done
