"""Many chains of ONE locus picked by its number of candidates: a single launch of identical units that saturates the GPU."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from bayesembler_b200 import synth  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=4)
ap.add_argument("--chains", type=int, default=32768)
ap.add_argument("--iters", type=int, default=300)
args = ap.parse_args()
b = synth.SynthBatch(2, n_loci=2000)
cand = [i for i in range(b.n) if b.T[i] == args.T and 20 <= b.R[i] <= 60]
i = cand[len(cand) // 2]
d = b.desc_array(n_chains=args.chains, subset=[i])
d["burn_in"] = args.iters // 11
d["samples"] = args.iters - args.iters // 11
ctx = bb.GibbsContext(0)
ctx.set_calibration(0)
ctx.submit(d); ctx.upload(); ctx.run(); ctx.sync()
t0 = time.perf_counter(); ctx.run(); ctx.sync(); dt = time.perf_counter() - t0
shape, clock = ctx.unit_stats()
dur = (clock[:, 1] - clock[:, 0]) * 1e-6
print("locus %d T=%d R=%d cells=%d: %d units, shape0 %s, launches %d, wall %.1f ms, %.1f M chain-iter/s, unit us/iter %.1f" % (
    i, b.T[i], b.R[i], b.cell_off[i + 1] - b.cell_off[i], len(shape), shape[0, :7].tolist(), ctx.last_launches(), dt * 1e3,
    args.chains * args.iters / dt / 1e6, 1e3 * dur.mean() / args.iters))
