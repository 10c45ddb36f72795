set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
python tools/schedule_report.py --top 6 > gpurun_out/sched_r2e_fast.txt 2>&1; head -42 gpurun_out/sched_r2e_fast.txt
for B in 8000 16000 24000; do echo "== budget $B"; BAYES_UNIT_SMEM=$B python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps"; done
echo "== carveout 100"; BAYES_CARVEOUT=100 python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps"
