set -u
mkdir -p gpurun_out
for B in 7000 10000 14000; do
echo "== budget $B, 1 warp unless critical, carveout 100"; BAYES_UNIT_SMEM=$B BAYES_SMEM_PER_WARP=60000 BAYES_CARVEOUT=100 python tools/schedule_report.py --top 3 2>&1 | grep -E "wall clock|makespan|warp-time|n=|resident warps|^ +[0-9]+ \|" | head -24
done
