set -u
mkdir -p gpurun_out
for B in 6000 9000 14000; do for W in 1; do for C in 65 100; do
echo "== budget $B warps $W carveout $C"; BAYES_UNIT_SMEM=$B BAYES_FORCE_WARPS=$W BAYES_CARVEOUT=$C python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time"
done; done; done
echo "== budget 9000, model-chosen warps, carveout 100"; BAYES_UNIT_SMEM=9000 BAYES_CARVEOUT=100 python tools/schedule_report.py --top 4 2>&1 | grep -vE "^ +[0-9]+ \|" | head -30
