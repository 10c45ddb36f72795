#!/bin/bash
# A/B timing of a compile-time switch: bash tools/exp_ab.sh "-DFLAG_A" "-DFLAG_B" ...
for F in "$@"; do
  BAYES_NVCC_FLAGS="$F" python -c "from bayesembler_b200 import _build; _build.build(force=True)"
  echo "=== flags: $F"
  python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time"
done
python -c "from bayesembler_b200 import _build; _build.build(force=True)"
