"""Saturated-regime run of a MIX of loci: the first N loci of a configuration, C chains each, a shortened chain -- the
units form as they do in a full batch (bundles of several slots), in a few launches that fill the GPU: the command to
profile with ncu when the full batch is too long to replay.

    python tools/bulk_mix.py [--config 2] [--loci 400] [--chains 16] [--iters 600]
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
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--loci", type=int, default=400)
ap.add_argument("--chains", type=int, default=16)
ap.add_argument("--iters", type=int, default=600)
ap.add_argument("--maxT", type=int, default=16)
args = ap.parse_args()
batch = synth.SynthBatch(args.config, n_loci=args.loci)
keep = np.nonzero((batch.T > 1) & (batch.T <= args.maxT))[0]
d = batch.desc_array(n_chains=args.chains, subset=keep)
d["burn_in"] = args.iters // 11
d["samples"] = args.iters - args.iters // 11
ctx = bb.GibbsContext(0)
ctx.set_calibration(0)
ctx.submit(d)
ctx.upload()
ctx.run()
ctx.sync()
t0 = time.perf_counter()
ctx.run()
ctx.sync()
dt = time.perf_counter() - t0
shape, clock = ctx.unit_stats()
dur = (clock[:, 1] - clock[:, 0]) * 1e-6
jobs = len(keep) * args.chains
print("loci %d x %d chains = %d jobs in %d units (%.1f slots/unit), %d launches, wall %.1f ms, %.1f M chain-iterations/s, unit us/iter mean %.1f" % (
    len(keep), args.chains, jobs, len(shape), jobs / len(shape), ctx.last_launches(), dt * 1e3, jobs * args.iters / dt / 1e6,
    1e3 * dur.mean() / args.iters))
by = {}
for s in shape:
    by[(int(s[1]), int(s[2]))] = by.get((int(s[1]), int(s[2])), 0) + 1
print("units by (threads, slots):", sorted(by.items()))
ctx.close()
