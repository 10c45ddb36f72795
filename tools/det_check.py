import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import bayesembler_b200 as bb
from oracle.oracle_py import random_locus
rng = np.random.default_rng(77)
shapes = [(1, 3, 9), (2, 6, 40), (5, 50, 100), (12, 70, 30), (40, 120, 200)]
loci = [random_locus(rng, T, R, max_count=mc) for (T, R, mc) in shapes]
seeds = [int(s) for s in rng.integers(1, 2 ** 31, len(loci))]
burn = [30 + 3 * l.T for l in loci]
samples = [10 * b for b in burn]
ctx = bb.GibbsContext(0)
ctx.run_batch(bb.make_desc_array(loci, burn, samples, seeds))
whole = [ctx.fetch(i) for i in range(len(loci))]
sh, _ = ctx.unit_stats(); print("batch unit shapes", sh.tolist())
ctx.clear()
ctx.run_batch(bb.make_desc_array(loci, burn, samples, seeds))
again = [ctx.fetch(i) for i in range(len(loci))]
for i in range(len(loci)):
    print("rerun", i, all(np.array_equal(whole[i][k], again[i][k]) for k in whole[i]))
for i in range(len(loci)):
    ctx.clear()
    ctx.run_batch(bb.make_desc_array([loci[i]], [burn[i]], [samples[i]], [seeds[i]]))
    one = ctx.fetch(0)
    sh, _ = ctx.unit_stats()
    print("alone", i, sh.tolist(), {k: bool(np.array_equal(whole[i][k], one[k])) for k in one})
    if not np.array_equal(whole[i]["mean"], one["mean"]):
        print(whole[i]["mean"] - one["mean"], whole[i]["occ"])
    # explicit pi
    pi = ctx.locus_info(0)["pi"]
    ctx.clear()
    ctx.run_batch(bb.make_desc_array([loci[i]], [burn[i]], [samples[i]], [seeds[i]], pi=[pi]))
    two = ctx.fetch(0)
    print("alone, pi given", i, {k: bool(np.array_equal(whole[i][k], two[k])) for k in two})
