"""Likelihood builder (row f1) timing, host CPU only: fragments/s of the reference's AlignmentParser (oracle/_ref, dense
matrix + linear row search), the C oracle and the product (sparse rows + hashed lookup) on the same synthetic loci,
one thread each, plus the product over a thread pool of loci (how the reference's workers would call it).

    python tools/bench_f1.py [--loci 60] [--frags 3000]
"""
import argparse
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from oracle import f1_py  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--loci", type=int, default=60)
ap.add_argument("--frags", type=int, default=3000)
args = ap.parse_args()
rng = np.random.default_rng(7)
loci = [f1_py.random_raw_locus(rng, n_exons=int(rng.integers(3, 12)), n_cand=int(rng.integers(2, 16)), n_frag=args.frags) for _ in range(args.loci)]
n_frag = sum(l.n_frag for l in loci)
model = bb.SequencingModel(200.0, 20.0, max(int(l.cand_length.max()) for l in loci))


def timed(fn):
    t0 = time.perf_counter()
    out = [fn(l) for l in loci]
    return time.perf_counter() - t0, out


res = {}
if os.path.exists(f1_py.REF_F1_SO):
    ref = f1_py.ReferenceF1()
    res["reference (unmodified TUs, 1 thread)"], want = timed(lambda l: ref.collapsed_map(l, max_len=model.max_len))
orc = f1_py.OracleF1()
res["C oracle (dense, linear search, 1 thread)"], o = timed(lambda l: orc.collapsed_map(l, max_len=model.max_len))
res["product (sparse, hashed, 1 thread)"], g = timed(lambda l: bb.build_collapsed_map(l, model))
n_threads = os.cpu_count() or 1
t0 = time.perf_counter()
with ThreadPoolExecutor(n_threads) as ex:
    list(ex.map(lambda l: bb.build_collapsed_map(l, model), loci))
res["product (%d threads over loci)" % n_threads] = time.perf_counter() - t0
rows = sum(x["R"] for x in g)
print("%d loci, %d fragments, %d collapsed rows" % (len(loci), n_frag, rows))
for k, v in res.items():
    print("  %-45s %8.3f s  %10.0f fragments/s" % (k, v, n_frag / v))
for a, b in zip(g, o):
    assert a["R"] == b["R"] and np.array_equal(a["counts"], b["counts"])
