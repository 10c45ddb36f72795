"""Run one synthetic locus (optionally many chains of it) once -- the smallest command worth profiling with ncu."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from bayesembler_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--locus", type=int, default=1789)
ap.add_argument("--chains", type=int, default=1)
ap.add_argument("--iters", type=int, default=2200)
args = ap.parse_args()
batch = synth.SynthBatch(args.config, ids=[args.locus])
d = batch.desc_array(n_chains=args.chains)
d["burn_in"] = args.iters // 11
d["samples"] = args.iters - args.iters // 11
ctx = bb.GibbsContext(0)
ctx.run_batch(d)
r = ctx.fetch(0)
print("T=%d R=%d occ=%s" % (batch.T[0], batch.R[0], r["occ"]))
ctx.close()
