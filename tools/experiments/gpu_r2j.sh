set -u
mkdir -p gpurun_out
TAG=r2j_mix1w
CMD="env BAYES_FORCE_WARPS=1 python tools/bulk_mix.py --loci 400 --chains 64 --iters 400"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -s 0 -c 3 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log
BAYES_FORCE_WARPS=1 BAYES_MODE=replay python tools/bulk_mix.py --loci 400 --chains 64 --iters 400
BAYES_MODE=replay python tools/bulk_mix.py --loci 400 --chains 64 --iters 400
BAYES_FORCE_WARPS=1 BAYES_TILE_STREAM=0 python tools/bulk_mix.py --loci 400 --chains 64 --iters 400
BAYES_FORCE_WARPS=1 BAYES_UNIT_SMEM=20000 python tools/bulk_mix.py --loci 400 --chains 64 --iters 400
BAYES_FORCE_WARPS=2 BAYES_UNIT_SMEM=20000 python tools/bulk_mix.py --loci 400 --chains 64 --iters 400
