"""Candidate enumeration (row f2) timing, host CPU only: graphs/s and candidates/s of the reference's AllPaths
(oracle/_ref, Boost.Graph stand-in, recursive enumeration restarted per junction-threshold step) and the product
(path-count dynamic programme + one iterative enumeration) on the same synthetic splice graphs, one thread each.

    python tools/bench_f2.py [--graphs 40] [--exons 40] [--max-candidates 1000]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from oracle import f2_py  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--graphs", type=int, default=40)
ap.add_argument("--exons", type=int, default=40)
ap.add_argument("--max-candidates", type=int, default=1000)
ap.add_argument("--p-edge", type=float, default=0.6)
args = ap.parse_args()
dots = [f2_py.random_splice_graph(1000 + i, args.exons, p_edge=args.p_edge, max_cov=200) for i in range(args.graphs)]


def timed(fn):
    t0 = time.perf_counter()
    out = [fn(d) for d in dots]
    return time.perf_counter() - t0, out


res = {}
want = None
if os.path.exists(f2_py.REF_F2_SO):
    ref = f2_py.ReferenceF2()
    res["reference (unmodified allPaths.cpp, 1 thread)"], want = timed(lambda d: ref.all_paths(d, False, args.max_candidates))
res["product (1 thread)"], got = timed(lambda d: bb.all_paths(d, False, args.max_candidates))
if want is not None:
    for w, g in zip(want, got):
        assert w["names"] == g["names"] and np.array_equal(w["exon_start"], g["exon_start"]) and np.array_equal(w["exon_end"], g["exon_end"])
n_cand = sum(g["n_cand"] for g in got)
adj = np.array([g["stats"][3] for g in got])
print("%d graphs x %d exons, --max-candidate-number %d: %d candidates, junction-threshold adjustments mean %.1f max %d"
      % (len(dots), args.exons, args.max_candidates, n_cand, adj.mean(), adj.max()))
for k, v in res.items():
    print("  %-50s %8.3f s  %9.1f graphs/s  %10.0f candidates/s" % (k, v, len(dots) / v, n_cand / v))
