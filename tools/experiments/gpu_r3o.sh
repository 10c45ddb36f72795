set -u
mkdir -p gpurun_out
for V in 128 32 128 32; do
  if [ "$V" = "128" ]; then unset BAYES_B200_LIB; else export BAYES_B200_LIB=$PWD/bayesembler_b200/libbayes_chunks32.so; fi
  python tools/schedule_report.py --config 4 --top 1 2>&1 | grep -E "wall clock" | sed "s/^/chunks $V: /"
done
