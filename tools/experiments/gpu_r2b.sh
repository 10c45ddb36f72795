set -u
mkdir -p gpurun_out
bash tools/ncu_bulk.sh r2a_fast_bulk
BAYES_MODE=replay bash tools/ncu_bulk.sh r2a_replay_bulk
python tools/bulk_shape.py 1220 4096 1100
BAYES_MODE=replay python tools/bulk_shape.py 1220 4096 1100
for W in 1 2 4; do echo "== fast forced warps $W"; BAYES_FORCE_WARPS=$W python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time"; done
