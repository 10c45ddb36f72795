#!/bin/bash
# Saturated-regime profile of the sampler kernel: many chains of one mid-size locus, every unit forced to the same shape
# (2 warps, 32-row slabs), = ONE launch that fills the GPU.
# usage: bash tools/ncu_bulk.sh <tag> [locus] [chains] [iters]
TAG=${1:-bulk}; LOCUS=${2:-1220}; CHAINS=${3:-4096}; ITERS=${4:-1100}
CMD="env BAYES_FORCE_WARPS=2 BAYES_FORCE_H=32 python tools/run_one.py --locus $LOCUS --chains $CHAINS --iters $ITERS"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -c 1 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
