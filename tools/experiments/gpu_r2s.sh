set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2s_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2s_bench_ref.json 2> gpurun_out/r2s_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2s_bench_ref.json | cut -c1-700
nproc
