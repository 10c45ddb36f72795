set -u
mkdir -p gpurun_out
for CR in 0.75 1.0 1.3; do echo "== 5000 loci, calibration on, crit $CR"; BAYES_CRIT=$CR python tools/schedule_report.py --config 4 --loci 5000 --top 3 > gpurun_out/sched_r3l_$CR.txt 2>&1
grep -E "wall clock|makespan|resident warps/SM over" gpurun_out/sched_r3l_$CR.txt; done
echo "== 2500 loci"; python tools/schedule_report.py --config 4 --loci 2500 --top 3 2>&1 | grep -E "wall clock|makespan|resident warps/SM over"
echo "== 2500 loci, no calibration"; BAYES_CALIB_ITERS=0 python tools/schedule_report.py --config 4 --loci 2500 --top 3 2>&1 | grep -E "wall clock|makespan|resident warps/SM over"
