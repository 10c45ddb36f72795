#!/bin/bash
# timing-only experiment: BINV/BTRD switch point (parity needs the oracle rebuilt with the same value)
for NQ in 10.0 16.0 24.0 32.0 48.0; do
  BAYES_NVCC_FLAGS="-DBAYES_BINV_NQ=$NQ" python -c "from bayesembler_b200 import _build; _build.build(force=True)"
  echo "=== BINV_NQ=$NQ"
  python tools/schedule_report.py --top 3 2>&1 | grep -E "makespan|warp-time"
  python tools/locus_timing.py --chains 1,4096 --T 5,12,30 2>&1 | tail -6
done
python -c "from bayesembler_b200 import _build; _build.build(force=True)"
