set -u
mkdir -p gpurun_out
for CAL in 0 256 1024; do echo "== 5000 loci, calibration $CAL"; BAYES_CALIB_ITERS=$CAL python tools/schedule_report.py --config 4 --loci 5000 --top 3 > gpurun_out/sched_r3k_$CAL.txt 2>&1
grep -E "wall clock|makespan|^  \(0|resident warps/SM over|time model" gpurun_out/sched_r3k_$CAL.txt; sed -n '/^longest/,/^first in launch/p' gpurun_out/sched_r3k_$CAL.txt | head -5; done
