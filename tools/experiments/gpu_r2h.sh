set -u
mkdir -p gpurun_out
TAG=r2g_slice
CMD="python tools/schedule_report.py --loci 4000 --top 1"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -s 40 -c 6 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log; grep -E "wall clock|makespan|warp-time|launches" gpurun_out/${TAG}_plain.log
