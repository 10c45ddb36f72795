set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
python tools/schedule_report.py --top 8 > gpurun_out/sched_r2a_fast.txt 2>&1; head -30 gpurun_out/sched_r2a_fast.txt
python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2a_bench_fast.json 2> gpurun_out/r2a_bench_fast.err; cat gpurun_out/r2a_bench_fast.json | cut -c1-600
BAYES_MODE=replay python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2a_bench_replay.json 2> gpurun_out/r2a_bench_replay.err; cut -c1-400 gpurun_out/r2a_bench_replay.json
