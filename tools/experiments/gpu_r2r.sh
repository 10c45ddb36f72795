set -u
python tools/schedule_report.py --config 2 --top 3 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps|n="
python tools/schedule_report.py --config 4 --top 3 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps|n="
