set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python tools/schedule_report.py --config 4 --top 6 > gpurun_out/sched_r2q_c4.txt 2>&1; head -30 gpurun_out/sched_r2q_c4.txt
python tools/schedule_report.py --config 2 --top 3 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps|n="
