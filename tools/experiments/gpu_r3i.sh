set -u
mkdir -p gpurun_out
for L in 5000 10000; do echo "== $L loci"; python tools/schedule_report.py --config 4 --loci $L --top 3 > gpurun_out/sched_r3i_$L.txt 2>&1
grep -E "wall clock|makespan|^  \(0|resident warps/SM over" gpurun_out/sched_r3i_$L.txt; sed -n '/^longest/,/^first in launch/p' gpurun_out/sched_r3i_$L.txt | head -5; done
