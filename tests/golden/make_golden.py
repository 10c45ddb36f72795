#!/usr/bin/env python
"""Generate tests/golden/ref_vectors.npz from the UNMODIFIED reference sampler (oracle/_ref, built from /root/reference/src
against oracle/shim by oracle/Makefile).  The reference ships no golden vectors of its own, so these are
outputs of the reference's code run in the build container; the GPU box has no /root/reference and checks the oracle
against this file instead.

    python tests/golden/make_golden.py        # needs /root/reference (or a prebuilt oracle/_ref/libbayesref.so)

Contents (all produced by reference code on a sequential MT19937, seeds below):
  case{i}_*      a small CollapsedMap (CSR) with: min read cover size, initAssignment counts, generateAssignment counts for
                 a fixed theta, generateExpression theta/plus order, the full runSampler accumulators (occ / mean / M2),
                 pi, and the number of 32-bit RNG words each call consumed
  table{j}_*     FixedBinomialFixedGammaSimplexSizeGenerator CDF tables
  top_*          getTopMarginals on fixed accumulators
  gtf{k}_*       writeEnsembleGtf / writeCandidateGtf text for oracle_py.gtf_case(100 + k, ...)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")

# (T, R, max_count, gamma, density, burn, samples)
CASES = [
    (1, 4, 30, 1.0, 0.4, 20, 200),
    (2, 6, 40, 1.0, 0.6, 50, 500),
    (3, 25, 60, 1.0, 0.5, 60, 600),
    (5, 40, 200, 1.0, 0.4, 80, 800),
    (8, 90, 15, 0.5, 0.4, 80, 800),       # every row below the 20-fragment threshold: CDF walk only
    (12, 60, 3000, 2.0, 0.3, 100, 1000),   # large counts: conditional binomial chain (BTRD)
    (30, 200, 80, 1.0, 0.25, 100, 1000),
    (40, 120, 400, 1.0, 0.15, 60, 600),    # T > 32: wide-unit path on the GPU
    (100, 250, 50, 1.0, 0.08, 40, 400),
]
TABLES = [(0.5, 1.0, 2, 10), (0.25, 1.0, 8, 500), (1.0 - 2.220446049250313e-16, 1.0, 5, 100), (0.1, 0.5, 30, 2000),
          (0.03, 2.0, 100, 9000), (0.6, 1.0, 64, 40)]


def main():
    oracle_py.build(ref=True)
    ref = oracle_py.Reference()
    out = {}
    meta = []
    for ci, (T, R, mc, gamma, dens, burn, samples) in enumerate(CASES):
        rng = np.random.default_rng(1000 + ci)
        loc = oracle_py.random_locus(rng, T, R, max_count=mc, density=dens, gamma=gamma)
        seed = int(rng.integers(1, 2 ** 31 - 1))
        p = "case%d_" % ci
        out[p + "row_ptr"], out[p + "col_idx"], out[p + "val"] = loc.row_ptr, loc.col_idx, loc.val
        out[p + "row_count"], out[p + "efflen"] = loc.row_count, loc.efflen
        meta.append((T, R, gamma, loc.n_total, seed, burn, samples))
        cov, w = ref.min_read_cover(loc, seed)
        out[p + "cover"] = np.array([cov, w], np.int64)
        c, w = ref.init_assignment(loc, seed)
        out[p + "init_counts"], out[p + "init_words"] = c, np.int64(w)
        if T > 1:
            th = rng.dirichlet(np.ones(T) * 0.7)
            if T > 3:
                th[rng.choice(T, size=T // 3, replace=False)] = 0.0
            P = loc.dense()
            for i in np.nonzero((P * th).sum(1) <= 0)[0]:
                th[loc.col_idx[loc.row_ptr[i]]] = rng.random() + 0.01
            th = th / th.sum()
            out[p + "theta_in"] = th
            c, w = ref.generate_assignment(loc, th, seed)
            out[p + "assign_counts"], out[p + "assign_words"] = c, np.int64(w)
            n_plus = int((c > 0).sum())
            b = min(T, n_plus + (T - n_plus + 1) // 2)
            e, order, w = ref.generate_expression(c, loc.n_frag, b, gamma, seed)
            out[p + "expr_b"], out[p + "expr_theta"], out[p + "expr_order"], out[p + "expr_words"] = np.int64(b), e, order, np.int64(w)
        r = ref.run_locus(loc, burn, samples, seed)
        out[p + "occ"], out[p + "mean"], out[p + "m2"] = r["occ"], r["mean"], r["m2"]
        out[p + "run"] = np.array([r["cover"], r["words"]], np.int64)
        out[p + "pi"] = np.float64(r["pi"])
    out["meta"] = np.array(meta, np.float64)
    for ti, (pi, gamma, T, nf) in enumerate(TABLES):
        tab, rl = ref.simplex_table(pi, gamma, T, nf)
        out["table%d_val" % ti], out["table%d_len" % ti] = tab, rl
    out["tables"] = np.array(TABLES, np.float64)
    occ = np.array([100, 49, 50, 1, 0, 77], np.int32)
    mean = np.array([10.0, 3.5, 0.25, 9.0, 0.0, 1e-3])
    m2 = np.array([400.0, 12.0, 0.5, 0.0, 0.0, 1e-4])
    idx, mo, sd = ref.top_marginals(100, occ, mean, m2, 0.5)
    out["top_occ"], out["top_mean"], out["top_m2"] = occ, mean, m2
    out["top_idx"], out["top_mean_out"], out["top_sd_out"] = idx, mo, sd
    for gi in range(4):
        case = oracle_py.gtf_case(100 + gi, T=[7, 3, 12, 1][gi], samples=[40, 25, 60, 10][gi], n_cand=[9, 3, 20, 2][gi])
        out["gtf%d_ensemble" % gi] = np.frombuffer(ref.gtf(case, 0).encode(), np.uint8)
        out["gtf%d_candidates" % gi] = np.frombuffer(ref.gtf(case, 1).encode(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %d bytes)" % (OUT, len(out), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
