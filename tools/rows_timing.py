"""Per-iteration time of config-3 class loci (more than 32 candidates, gibbs_rows.cu) against the cluster size.

    python tools/rows_timing.py [--loci 3] [--iters 3000] [--ctas 1,2,4,8,16]

For a few loci of config 3 spread over the candidate-count order (always including the largest): time per Gibbs iteration
of the locus running alone, for every cluster size, and the SM-time it costs (CTAs x time: what throughput pays for latency).
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
ap.add_argument("--loci", type=int, default=3)
ap.add_argument("--iters", type=int, default=3000)
ap.add_argument("--ctas", default="1,2,4,8,16")
ap.add_argument("--mode", default="fast")
ap.add_argument("--pick", type=int, default=-1, help="one locus id instead of the spread")
ap.add_argument("--together", action="store_true", help="also run the picked loci as one batch per cluster size")
args = ap.parse_args()
batch = synth.SynthBatch(3)
order = np.argsort(batch.T)
picks = [int(order[-1])] + [int(order[(2 * k + 1) * len(order) // (2 * max(1, args.loci - 1))]) for k in range(args.loci - 1)]
if args.pick >= 0:
    picks = [args.pick]
nnz = batch.cell_off[1:] - batch.cell_off[:-1]
print("%6s %5s %5s %7s %6s %4s | %10s %12s" % ("locus", "T", "R", "nnz", "N_f", "CTAs", "us/iter", "SM-us/iter"))
for i in picks:
    for nc in [int(c) for c in args.ctas.split(",")]:
        os.environ["BAYES_FORCE_CTAS"] = str(nc)
        d = batch.desc_array(subset=[i])
        d["burn_in"] = args.iters // 11
        d["samples"] = args.iters - args.iters // 11
        ctx = bb.GibbsContext(0, mode=args.mode)
        ctx.submit(d)
        ctx.upload()
        ctx.run()
        ctx.sync()
        t0 = time.perf_counter()
        ctx.run()
        ctx.sync()
        dt = time.perf_counter() - t0
        ctx.close()
        us = 1e6 * dt / args.iters
        print("%6d %5d %5d %7d %6d %4d | %10.2f %12.1f" % (i, batch.T[i], batch.R[i], nnz[i], batch.n_frag[i], nc, us, us * nc), flush=True)
if args.together:
    for nc in [int(c) for c in args.ctas.split(",")]:
        os.environ["BAYES_FORCE_CTAS"] = str(nc)
        d = batch.desc_array(subset=picks)
        d["burn_in"] = args.iters // 11
        d["samples"] = args.iters - args.iters // 11
        ctx = bb.GibbsContext(0, mode=args.mode)
        ctx.submit(d)
        ctx.upload()
        ctx.run()
        ctx.sync()
        t0 = time.perf_counter()
        ctx.run()
        ctx.sync()
        dt = time.perf_counter() - t0
        ctx.close()
        print("batch of %d loci, %d CTAs each: %.1f ms for %d iterations" % (len(picks), nc, 1e3 * dt, args.iters), flush=True)
