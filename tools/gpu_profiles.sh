#!/bin/bash
# ncu evidence of the final code state (one B200): launch list of one bench step + `--set full` of the dominant kernel in the saturated regime.
set -u
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 180 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --no-cpu-baseline --no-other-configs --steps 1 --warmup 3 > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
CMD="env BAYES_FORCE_WARPS=1 python tools/bulk_one.py --T 4 --chains 65536 --iters 200"
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -s 0 -c 1 -f -o $OUT/${TAG}_sat_prof $CMD > $OUT/${TAG}_sat_ncu.log 2>&1
echo "ncu saturated rc=$?"
