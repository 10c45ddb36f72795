"""The whole host + device chain at scale through the command line: N synthetic splice graphs with fragment alignments ->
graph file -> bayesembler_b200_cli (AllPaths -> AlignmentParser -> one GPU sampler batch -> GTF) with its own timing line.

    python tools/cli_pipeline.py [--graphs 300] [--threads 16]
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bayesembler_b200 import _build, synth  # noqa: E402
import test_f3_cli as t3  # noqa: E402  (graph and fragment generators of the CLI tests)

ap = argparse.ArgumentParser()
ap.add_argument("--graphs", type=int, default=300)
ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
ap.add_argument("--frags", type=int, default=1000)
args = ap.parse_args()
cli = _build.build_cli()
rng = np.random.default_rng(42)
t0 = time.perf_counter()
graphs = []
for g in range(args.graphs):
    strand = "+" if g % 2 == 0 else "-"
    dot = t3._long_exon_graph(1000 + g, 3 + g % 7, 0.7, "graph_%d_%s" % (g, "plus" if strand == "+" else "minus"), strand, pos=5000 + 40000 * g)
    graphs.append(([dot], t3._fragments(rng, dot, int(args.frags * (0.3 + 1.4 * rng.random()))), False))
n_pairs = sum(len(g[1].frag_prob) for g in graphs)
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "graphs.txt")
    synth.write_graph_file(path, graphs, n_pairs)
    print("generated %d graphs, %d fragments, %.1f MB graph file in %.1f s (Python)" % (len(graphs), n_pairs, os.path.getsize(path) / 1e6,
                                                                                        time.perf_counter() - t0))
    for threads in sorted({1, args.threads}):
        t1 = time.perf_counter()
        p = subprocess.run([cli, "-g", path, "-o", os.path.join(tmp, "out_%d_" % threads), "--seed", "1", "--frag-mean", "200", "--frag-sd", "20",
                            "--num-threads", str(threads)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        wall = time.perf_counter() - t1
        if p.returncode != 0:
            print(p.stderr)
            sys.exit(1)
        lines = p.stdout.strip().split("\n")
        print("--num-threads %d: wall %.2f s (%.0f graphs/s)\n  %s\n  %s" % (threads, wall, len(graphs) / wall,
              [l for l in lines if "assembled successfully" in l][0].strip(), lines[-1]))
        n_tx = sum(1 for l in open(os.path.join(tmp, "out_%d_assembly.gtf" % threads)) if l.strip())
        print("  assembly.gtf: %d exon lines" % n_tx)
