set -u
export BAYES_B200_LIB=$PWD/bayesembler_b200/libbayes_phases.so
for L in 11 459; do for N in 1 4 16; do BAYES_FORCE_CTAS=$N python tools/run_one.py --config 3 --locus $L --iters 3000 | grep "rows phases"; done; done
