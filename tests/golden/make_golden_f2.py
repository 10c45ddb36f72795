#!/usr/bin/env python
"""Generate tests/golden/f2_vectors.json from the UNMODIFIED reference candidate enumeration (oracle/_ref/libbayesref_f2.so:
allPaths.cpp compiled against oracle/shim/orc_graph_shim.hpp by oracle/Makefile).

    python tests/golden/make_golden_f2.py        # needs /root/reference (or a prebuilt oracle/_ref/libbayesref_f2.so)

Every case = the arguments of oracle.f2_py.random_splice_graph (or a literal DOT text), (no_pre_mrna,
max_candidate_number), and the reference's answer: candidate names, flat exon arrays, lengths, pre-mRNA flags and the
numbers of its log block (vertices, sources, sinks, junction-threshold adjustments, paths).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import f2_py  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "f2_vectors.json")
# seed, n_exons, p_edge, pre_mrna, no_pre_mrna, max_candidate_number
CASES = [(1, 2, 0.9, True, False, 100), (2, 3, 0.9, True, False, 100), (3, 5, 0.5, True, True, 100), (4, 8, 0.6, True, False, 100),
         (5, 12, 0.5, True, False, 100), (6, 12, 0.7, True, False, 10), (7, 14, 0.8, False, False, 25), (8, 16, 0.6, True, True, 5),
         (9, 11, 0.9, True, False, 3), (10, 13, 0.3, True, False, 2), (11, 9, 0.2, False, True, 1), (12, 15, 0.9, True, False, 0),
         (13, 18, 0.7, True, False, 60)]
LITERAL = [
    # one exon and an identical pre-mRNA: the pre-mRNA vertex is dropped (allPaths.cpp:136-167)
    ('digraph graph_1_minus {\n0 [label="chrX;-;500;900"]\npre_mrna [label="chrX;-;500;900"]\n}\n', False, 100),
    # same, but the pre-mRNA is longer: both stay
    ('digraph graph_2_minus {\n0 [label="chrX;-;500;900"]\npre_mrna [label="chrX;-;400;900"]\n}\n', False, 100),
    # parallel junction statements give the path twice
    ('digraph graph_3_plus {\n0 [label="c;+;1;10"]\n1 [label="c;+;11;20"]\n0->1 [coverage="4"]\n0->1 [coverage="6"]\n}\n', True, 100),
]


def cases():
    for c in CASES:
        seed, n_exons, p_edge, pre, nopre, maxc = c
        yield dict(gen=list(c)), f2_py.random_splice_graph(seed, n_exons, p_edge=p_edge, pre_mrna=pre), nopre, maxc
    for dot, nopre, maxc in LITERAL:
        yield dict(dot=dot, no_pre_mrna=nopre, max_candidate_number=maxc), dot, nopre, maxc


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    ref = f2_py.ReferenceF2()
    out = []
    for desc, dot, nopre, maxc in cases():
        r = ref.all_paths(dot, nopre, maxc)
        desc["want"] = dict(names=r["names"], graph_name=r["graph_name"], stats=[int(x) for x in r["stats"]],
                            **dict((k, [int(x) for x in r[k]]) for k in ("exon_ptr", "exon_start", "exon_end", "cand_length", "is_premrna")))
        out.append(desc)
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %s: %d cases, %d bytes" % (OUT, len(out), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
