#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of the same bench command and one
# `--set full` capture of the dominant kernel.  Run from the repo root under gpurun:
#     gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r1'
# Outputs land in gpurun_out/ (scratch); tools/summarise_profile.py turns them into profiles/<tag>_*.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.csv 2>&1

python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log

python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log

python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
rc=$?
echo "bench rc=$rc"
cat $OUT/${TAG}_bench.json
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "bench ref rc=$?"
cat $OUT/${TAG}_bench_ref.json

if [ $rc -eq 0 ]; then
  # launch list of the same command (cold-cache, serialised: compare shares, not absolutes) with the DRAM bytes of
  # every launch; the first 400 launches = two full steps of the warm-up, the rest of the run is not profiled
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
      --log-file $OUT/${TAG}_launches.csv python bench.py --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
fi

# the dominant kernel, full set: the first launch classes of a 1000-locus slice of the same workload ...
SMALL="python bench.py --loci 1000 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > $OUT/${TAG}_small.json 2> $OUT/${TAG}_small.err && \
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -c 3 -f -o $OUT/${TAG}_prof \
    $SMALL > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
# ... and the saturated regime: 4096 chains of one mid-size locus = ONE launch class that fills every SM
bash tools/ncu_bulk.sh ${TAG}_bulk 1220 4096 1100

# the other configurations at shortened chains (rates only; bench.py measures configs[1])
python tools/config_rate.py --config 3 --loci 444 --burn 20 --samples 200 > $OUT/${TAG}_configs.txt 2>&1
python tools/config_rate.py --config 4 --loci 40000 --burn 100 --samples 1000 >> $OUT/${TAG}_configs.txt 2>&1
python tools/config_rate.py --config 5 --loci 200 --chains 64 --burn 100 --samples 1000 >> $OUT/${TAG}_configs.txt 2>&1
python tools/config_rate.py --config 1 --loci 200 --burn 100 --samples 1000 >> $OUT/${TAG}_configs.txt 2>&1
cat $OUT/${TAG}_configs.txt
python tools/schedule_report.py > $OUT/${TAG}_sched.txt 2>&1
