"""Throughput of the BAM front end (row f4, host C++) on the build container's CPU: a synthetic coordinate-sorted paired-end BAM ->
duplicate removal + strand split (bayes_bam_front_open) -> instance file -> graphs with their fragments (bayes_bam_front_graphs).
The plain-Python restatement (oracle/f4_py.py) is timed beside it for scale; no reference binary of this stage can be built here.

    python tools/bench_f4.py [--loci 400] [--pairs 500]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from oracle import f4_py  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--loci", type=int, default=400)
ap.add_argument("--pairs", type=int, default=500)
ap.add_argument("--python-too", action="store_true")
args = ap.parse_args()
rng = np.random.default_rng(1)
refs = tuple(("chr%d" % (i + 1), 30000 * (args.loci // 4 + 2)) for i in range(4))
t0 = time.perf_counter()
refs, recs, instance, loci = f4_py.synth_dataset(rng, n_loci=args.loci, pairs_per_locus=args.pairs, refs=refs)
tmp = tempfile.mkdtemp()
bam, inst, out = os.path.join(tmp, "lib.bam"), os.path.join(tmp, "lib.instance"), os.path.join(tmp, "graphs.txt")
f4_py.write_bam(bam, refs, recs)
open(inst, "w").write(instance)
print("synthetic library: %d alignments, %d loci, BAM %.1f MB (written by the Python fixture writer in %.1f s)" % (
    len(recs), len(loci), os.path.getsize(bam) / 1e6, time.perf_counter() - t0))
t0 = time.perf_counter()
front = bb.BamFront(bam, "unstranded")
t1 = time.perf_counter()
n = front.graphs(inst, out)
t2 = time.perf_counter()
c = front.counts()
front.close()
print("product: read + filter + duplicate removal %.3f s (%.2f M alignments/s); instance -> %d graphs with fragments (%d read pairs kept, "
      "graph file %.1f MB) %.3f s" % (t1 - t0, len(recs) / (t1 - t0) / 1e6, n["with_fragments"], int(c["pairs_written"]),
                                      os.path.getsize(out) / 1e6, t2 - t1))
if args.python_too:
    t0 = time.perf_counter()
    _, parsed = f4_py.read_bam(bam)
    (written,), counts = f4_py.dedup(parsed, "unstranded")
    t1 = time.perf_counter()
    unsorted, _, _ = f4_py.parse_instance(instance, ".", 0.05, False)
    graphs = f4_py.collapse(unsorted, False)
    k = 0
    for g in graphs[:40]:
        k += len(f4_py.fetch(written, [r[0] for r in refs].index(g["reference"]), g))
    t2 = time.perf_counter()
    print("Python restatement: read + duplicate removal %.1f s (%.3f M alignments/s); fetch of the first 40 graphs %.1f s" % (
        t1 - t0, len(recs) / (t1 - t0) / 1e6, t2 - t1))
    assert counts == c
