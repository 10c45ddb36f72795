set -u
mkdir -p gpurun_out
TAG=r2i_mix
CMD="python tools/bulk_mix.py --loci 400 --chains 16 --iters 600"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -s 0 -c 4 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log
BAYES_MODE=replay python tools/bulk_mix.py --loci 400 --chains 16 --iters 600
for B in 4000 10000 20000; do echo "== budget $B"; BAYES_UNIT_SMEM=$B python tools/bulk_mix.py --loci 400 --chains 16 --iters 600; done
