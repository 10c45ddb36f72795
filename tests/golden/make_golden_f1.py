#!/usr/bin/env python
"""Generate tests/golden/f1_vectors.npz from the UNMODIFIED reference likelihood builder (oracle/_ref/libbayesref_f1.so:
alignmentParser.cpp, sequencingModel.cpp, fragmentLengthModel.cpp compiled against oracle/shim by oracle/Makefile).

    python tests/golden/make_golden_f1.py        # needs /root/reference (or a prebuilt oracle/_ref/libbayesref_f1.so)

case{i}_*: a synthetic locus (candidates + paired-end alignments, oracle.f1_py.random_raw_locus with the arguments in
CASES) and the reference's CollapsedMap for it: return code, dense probability matrix, counts, new -> old candidate
index, effective lengths, decremented sub-graph complexities, fragment length histogram of fetchFragmentLengths.
tables_* / model_*: SequencingModel tables and the median / MAD fragment length estimate.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import f1_py  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "f1_vectors.npz")
# seed, n_exons, n_cand, n_frag, strand_plus, two_subgraphs, no_coverage_exclude
CASES = [(1, 2, 1, 40, True, False, False), (2, 4, 3, 200, False, False, False), (3, 6, 5, 400, True, False, False),
         (4, 6, 5, 400, False, True, False), (5, 8, 8, 600, True, True, False), (6, 5, 4, 300, True, False, True),
         (7, 3, 2, 1500, False, False, False), (8, 9, 8, 900, True, False, False), (9, 6, 6, 25, True, False, False)]
FIELDS = ["exon_ptr", "exon_start", "exon_end", "cand_length", "is_premrna", "subgraph_id", "read_pos", "cigar_ptr", "cigar_op", "cigar_len",
          "frag_prob"]


def make_case(c):
    seed, n_exons, n_cand, n_frag, plus, two, nce = c
    rng = np.random.default_rng(5000 + seed)
    return f1_py.random_raw_locus(rng, n_exons=n_exons, n_cand=n_cand, n_frag=n_frag, strand_plus=plus, two_subgraphs=two), nce


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    ref = f1_py.ReferenceF1()
    out = {}
    for ci, c in enumerate(CASES):
        loc, nce = make_case(c)
        p = "case%d_" % ci
        for f in FIELDS:
            out[p + f] = getattr(loc, f)
        out[p + "strand_plus"] = np.int32(loc.strand_plus)
        r = ref.collapsed_map(loc, no_coverage_exclude=nce)
        for k in ("P", "counts", "new_to_old", "efflen", "complexity"):
            out[p + k] = r[k]
        out[p + "rc"] = np.array([r["return_code"], r["T"], r["R"], int(nce)], np.int32)
        h = ref.fragment_length_histogram(loc)
        out[p + "hist"] = np.array(sorted(h.items()), np.int64).reshape(-1, 2)
    t = ref.sequencing_tables(200.0, 20.0, 4000)
    for k in t:
        out["tables_" + k] = t[k]
    hist = {int(k): int(v) for k, v in zip(range(140, 270), np.random.default_rng(1).integers(1, 200, 130))}
    out["model_hist"] = np.array(sorted(hist.items()), np.int64)
    out["model_est"] = np.array(ref.estimate_fragment_model(hist))
    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %d bytes)" % (OUT, len(out), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
