"""Development aid: per-phase time of the largest bundle loci of a configuration running alone (needs a library built with
-DBAYES_TILE_PHASES, see BAYES_B200_LIB)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from bayesembler_b200 import synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
batch = synth.SynthBatch(cfg)
cost = (batch.cell_off[1:] - batch.cell_off[:-1]).astype(np.float64) * (batch.T > 1) * (batch.T <= 32)
top = np.argsort(-cost)[:3]
for i in top:
    nnz = int(batch.cell_off[i + 1] - batch.cell_off[i])
    r0, r1 = batch.row_off[i], batch.row_off[i + 1]
    cnt = batch.count[r0:r1]
    print("locus %d: T=%d R=%d nnz=%d N_f=%d rows with >= 20 fragments %d (max %d), full iterations %d" % (
        i, batch.T[i], batch.R[i], nnz, batch.n_frag[i], int((cnt >= 20).sum()), int(cnt.max()), int(batch.burn[i] + batch.samples[i])), flush=True)
    for warps in (256, 128, 64, 32, 8):
        os.environ["BAYES_FORCE_WARPS"] = str(warps)
        d = batch.desc_array(subset=[int(i)])
        d["burn_in"] = iters // 11
        d["samples"] = iters - iters // 11
        ctx = bb.GibbsContext(0)
        ctx.submit(d)
        ctx.upload()
        ctx.run()
        ctx.sync()
        t0 = time.perf_counter()
        ctx.run()
        ctx.sync()
        dt = time.perf_counter() - t0
        ctx.close()
        print("  %2d warps: %.1f us/iteration" % (warps, 1e6 * dt / iters), flush=True)
