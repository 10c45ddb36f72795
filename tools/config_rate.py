"""Rate of a synthetic configuration with SHORTENED chains (the per-iteration cost of a unit does not depend on how many
iterations follow): cell-visits/s, draws/s and the algorithmic GB/s of SURVEY.md 8d for configurations whose full
iteration counts take minutes per step (config 3: 200-1000 candidates per locus, 150 000-660 000 iterations, matrices
of 0.8-1.6 MB streamed from L2/HBM every iteration).

    python tools/config_rate.py --config 3 --loci 148 --burn 40 --samples 400
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from bayesembler_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--loci", type=int, default=148)
ap.add_argument("--chains", type=int, default=1)
ap.add_argument("--burn", type=int, default=40)
ap.add_argument("--samples", type=int, default=400)
ap.add_argument("--runs", type=int, default=2)
args = ap.parse_args()
batch = synth.SynthBatch(args.config, n_loci=args.loci)
full_iters = float((batch.burn.astype(np.int64) + batch.samples)[batch.T > 1].sum())
batch.burn[:] = args.burn
batch.samples[:] = args.samples
d = batch.desc_array(n_chains=args.chains)
st = batch.stats(n_chains=args.chains)
ctx = bb.GibbsContext(0)
ctx.submit(d)
ctx.upload()
ctx.run()
ctx.sync()
t0 = time.perf_counter()
for _ in range(args.runs):
    ctx.run()
ctx.sync()
dt = (time.perf_counter() - t0) / args.runs
shape, clock = ctx.unit_stats()
dur = (clock[:, 1] - clock[:, 0]).astype(np.float64) * 1e-3 / (args.burn + args.samples)
print("config %d: %d loci x %d chain(s), T %d-%d, %d units (wide %d, streaming %d), %d launches" % (
    args.config, args.loci, args.chains, batch.T.min(), batch.T.max(), len(shape), int(shape[:, 0].sum()), int(((shape[:, 7] & 1) == 0).sum()),
    ctx.last_launches()))
print("  %d + %d iterations per chain: %.1f ms per run -> %.3g cell-visits/s, %.3g draws/s, %.1f GB/s algorithmic (12 nnz + 8 R + 24 T per iteration)"
      % (args.burn, args.samples, dt * 1e3, st["cell_visits"] / dt, st["draws"] / dt, st["algo_bytes"] / dt / 1e9))
print("  per-iteration time of a unit: median %.1f us, max %.1f us; at the full iteration counts (%.3g iterations in total) the slowest unit "
      "would run %.0f s" % (np.median(dur), dur.max(), full_iters, (dur * 1e-6 * full_iters / max(1, len(dur))).max()))
ctx.close()
