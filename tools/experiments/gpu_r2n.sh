set -u
for K in 0.6 0.5 0.4 0.3; do echo "== crit $K"; BAYES_CRIT=$K python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps|n=|^ +[0-9]+ \|" | head -12; done
