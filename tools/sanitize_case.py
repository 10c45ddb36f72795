"""Small mixed batch for compute-sanitizer (memcheck / racecheck): bundles, a multi-warp unit, a wide unit and a
streaming wide unit, a few hundred iterations each."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from bayesembler_b200 import synth  # noqa: E402
from oracle.oracle_py import random_locus  # noqa: E402

rng = np.random.default_rng(3)
loci = [random_locus(rng, T, R, max_count=mc) for (T, R, mc) in [(2, 9, 40), (5, 60, 100), (12, 200, 300), (30, 400, 50), (64, 150, 60), (300, 120, 30)]]
ctx = bb.GibbsContext(0)
ctx.run_batch(bb.make_desc_array(loci, [20] * len(loci), [120] * len(loci), list(range(1, len(loci) + 1)), n_chains=[3, 2, 2, 1, 1, 1]))
out = ctx.fetch_all()
b3 = synth.SynthBatch(3, n_loci=1)
d = b3.desc_array()
d["burn_in"] = 3
d["samples"] = 12
ctx.clear()
ctx.run_batch(d)
shape, _ = ctx.unit_stats()
print("ok", int(out["occ"].sum()), "streaming unit:", bool((shape[:, 7] & 1).min() == 0))
ctx.close()
