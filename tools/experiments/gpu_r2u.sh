set -u
mkdir -p gpurun_out
echo "== 2000 iterations"; python tools/rows_timing.py --pick 11 --iters 2000 --ctas 1,4,16
echo "== 30000 iterations"; python tools/rows_timing.py --pick 11 --iters 30000 --ctas 1,4,16
for N in 1 8; do
TAG=r2u_rows_c${N}
CMD="env BAYES_FORCE_CTAS=$N python tools/run_one.py --config 3 --locus 11 --iters 12000"
ncu --set full --clock-control none --import-source on -k regex:gibbs_rows -s 0 -c 1 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log
done
