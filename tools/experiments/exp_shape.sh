#!/bin/bash
# timing experiment: warps x slab height for single loci running alone (latency) -- informs the unit time model
for cfg in "1 32" "2 32" "4 32" "8 32" "8 16" "16 16" "16 8" "32 8" "32 4"; do
  set -- $cfg
  echo "=== warps=$1 h=$2"
  BAYES_FORCE_WARPS=$1 BAYES_FORCE_H=$2 python tools/locus_timing.py --chains 1 --T 5,12,20,30 2>&1 | tail -4
done
