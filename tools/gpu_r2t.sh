set -u
for B in 8000 16000; do for C in 65; do echo "== c4 budget $B carveout $C"; BAYES_UNIT_SMEM=$B BAYES_CARVEOUT=$C python tools/schedule_report.py --config 4 --top 1 2>&1 | grep -E "makespan|warp-time|resident warps"; done; done
for C in 85 100; do echo "== c4 budget default carveout $C"; BAYES_CARVEOUT=$C python tools/schedule_report.py --config 4 --top 1 2>&1 | grep -E "makespan|warp-time|resident warps"; done
echo "== c4 crit 0.4"; BAYES_CRIT=0.4 python tools/schedule_report.py --config 4 --top 1 2>&1 | grep -E "makespan|warp-time|resident warps"
