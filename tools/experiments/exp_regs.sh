#!/bin/bash
for R in 64 72 80 96; do
  BAYES_NVCC_FLAGS="-DBAYES_MAXNREG=$R" python -c "from bayesembler_b200 import _build; _build.build(force=True)"
  echo "=== MAXNREG=$R"
  python tools/schedule_report.py --top 1 2>&1 | grep -E "wall clock|makespan|warp-time"
  python tools/bulk_shape.py 1220 4096 4000 | head -1
  python tools/bulk_shape.py 1789 8192 4000 | head -1
done
python -c "from bayesembler_b200 import _build; _build.build(force=True)"
