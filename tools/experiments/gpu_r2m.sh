set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/schedule_report.py --top 5 > gpurun_out/sched_r2m_fast.txt 2>&1; head -32 gpurun_out/sched_r2m_fast.txt
for C in 80 100; do echo "== carveout $C"; BAYES_CARVEOUT=$C python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps"; done
for B in 8000 20000; do echo "== budget $B"; BAYES_UNIT_SMEM=$B python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps"; done
