set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
for B in 5000 7168 10000 16000; do for C in 65 100; do echo "== stream, budget $B carveout $C"; BAYES_UNIT_SMEM=$B BAYES_CARVEOUT=$C python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps"; done; done
python tools/schedule_report.py --top 6 > gpurun_out/sched_r2g_fast.txt 2>&1; head -42 gpurun_out/sched_r2g_fast.txt
