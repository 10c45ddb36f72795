set -u
mkdir -p gpurun_out
for M in fast replay; do echo "=== mode $M"; 
BAYES_MODE=$M python tools/config_rate.py --config 3 --loci 148 --burn 500 --samples 4500 --runs 1 2>&1 | tail -3
done
python bench.py --no-other-configs --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2p_bench_c4.json 2> gpurun_out/r2p_bench_c4.err; echo "bench c4 rc=$?"; cut -c1-400 gpurun_out/r2p_bench_c4.json; tail -2 gpurun_out/r2p_bench_c4.err
BAYES_MODE=replay python bench.py --no-other-configs --no-cpu-baseline --steps 2 --warmup 1 2>/dev/null | cut -c1-250
python tools/schedule_report.py --config 4 --top 6 > gpurun_out/sched_r2p_c4.txt 2>&1; head -30 gpurun_out/sched_r2p_c4.txt
