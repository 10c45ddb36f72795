#!/bin/bash
# Final GPU-box visit of round 2 (one B200): parity tests, smoke, the bench line of both arms, the ncu launch list of one bench step,
# `--set full` captures of the two sampler kernels.  gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh r2z'
set -u
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.csv 2>&1
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log; tail -1 $OUT/${TAG}_smoke.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; rc=$?; echo "bench rc=$rc"; cut -c1-400 $OUT/${TAG}_bench.json
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 $OUT/${TAG}_bench_ref.json
# launch list of one step of the same workload (cold-cache, serialised under ncu: compare SHARES, not absolutes), with the DRAM bytes of every launch
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 180 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --no-cpu-baseline --no-other-configs --steps 1 --warmup 3 > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# the dominant kernel in the saturated regime: 65 536 chains of one small locus = one launch class that fills every SM
CMD="env BAYES_FORCE_WARPS=1 python tools/bulk_one.py --T 4 --chains 65536 --iters 200"
$CMD > $OUT/${TAG}_sat_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gibbs_ -s 0 -c 1 -f -o $OUT/${TAG}_sat_prof $CMD > $OUT/${TAG}_sat_ncu.log 2>&1
echo "ncu saturated rc=$?"
# the cooperative-rows kernel on the largest config-3 locus: one CTA (what a batch of hundreds of such loci gets) and a cluster of 16
for N in 1 16; do
  CMD="env BAYES_FORCE_CTAS=$N python tools/run_one.py --config 3 --locus 11 --iters 3000"
  ncu --set full --clock-control none --import-source on -k regex:gibbs_rows -s 0 -c 1 -f -o $OUT/${TAG}_rows_c${N}_prof $CMD > $OUT/${TAG}_rows_c${N}_ncu.log 2>&1
  echo "ncu rows c$N rc=$?"
done
python tools/rows_timing.py --loci 3 --iters 3000 --together > $OUT/${TAG}_rows_timing.txt 2>&1; cat $OUT/${TAG}_rows_timing.txt
python tools/schedule_report.py --config 4 --top 6 > $OUT/sched_${TAG}_c4.txt 2>&1; head -14 $OUT/sched_${TAG}_c4.txt
