set -u
mkdir -p gpurun_out
TAG=r2e_fast_bulk
CMD="env BAYES_FORCE_WARPS=1 python tools/run_one.py --locus 1220 --chains 4096 --iters 1100"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -c 1 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log
BAYES_FORCE_WARPS=1 python tools/bulk_shape.py 1220 4096 1100
