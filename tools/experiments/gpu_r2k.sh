set -u
mkdir -p gpurun_out
TAG=r2k_sat
CMD="env BAYES_FORCE_WARPS=1 python tools/bulk_one.py --T 4 --chains 65536 --iters 200"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -s 0 -c 1 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log
BAYES_FORCE_WARPS=1 BAYES_MODE=replay python tools/bulk_one.py --T 4 --chains 65536 --iters 200
BAYES_MODE=replay python tools/bulk_one.py --T 4 --chains 65536 --iters 200
BAYES_FORCE_WARPS=1 BAYES_UNIT_SMEM=3000 python tools/bulk_one.py --T 4 --chains 65536 --iters 200
BAYES_FORCE_WARPS=1 BAYES_UNIT_SMEM=30000 python tools/bulk_one.py --T 4 --chains 65536 --iters 200
