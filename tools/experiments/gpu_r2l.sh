set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
BAYES_FORCE_WARPS=1 python tools/bulk_one.py --T 4 --chains 65536 --iters 200
BAYES_FORCE_WARPS=1 BAYES_CARVEOUT=100 python tools/bulk_one.py --T 4 --chains 65536 --iters 200
python tools/schedule_report.py --top 4 > gpurun_out/sched_r2l_fast.txt 2>&1; head -30 gpurun_out/sched_r2l_fast.txt | grep -vE "^\s+[0-9]+ \|"
for B in 4000 12000; do echo "== budget $B"; BAYES_UNIT_SMEM=$B python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps"; done
echo "== carveout 100"; BAYES_CARVEOUT=100 python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time|resident warps"
TAG=r2l_sat
CMD="env BAYES_FORCE_WARPS=1 python tools/bulk_one.py --T 4 --chains 65536 --iters 200"
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -s 0 -c 1 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
