"""Likelihood / compatibility builder: candidates + paired-end alignments -> CollapsedMap.

Three implementations are compared: the UNMODIFIED reference translation units alignmentParser.cpp / sequencingModel.cpp /
fragmentLengthModel.cpp (oracle/_ref/libbayesref_f1.so, live where present, and recorded in tests/golden/f1_vectors.npz),
the plain-C restatement with the reference's dense matrix and linear row search (oracle/f1_oracle.c), and the product
(bayesembler_b200/csrc/likelihood.cpp: sparse rows, hashed row lookup) through the C ABI.  north_star asks for
bit-exact indices / compatibility sets and 1e-6 relative likelihoods; all three agree to the last bit.
"""
import os
import sys

import numpy as np
import pytest

from oracle import f1_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_f1  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "f1_vectors.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def orc():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return f1_py.OracleF1()


@pytest.fixture(scope="module")
def ref():
    from oracle import oracle_py
    oracle_py.build(ref=True)
    if not os.path.exists(f1_py.REF_F1_SO):
        pytest.skip("oracle/_ref/libbayesref_f1.so not built (no /root/reference here)")
    return f1_py.ReferenceF1()


def _dense(g):
    P = np.zeros((g["R"], g["T"]))
    for i in range(g["R"]):
        s, e = g["row_ptr"][i], g["row_ptr"][i + 1]
        P[i, g["col_idx"][s:e]] = g["val"][s:e]
        assert np.all(np.diff(g["col_idx"][s:e]) > 0) and np.all(g["val"][s:e] != 0)
    return P


def _same(a, b):
    assert a["return_code"] == b["return_code"]
    if a["return_code"] != 0:
        return
    assert a["T"] == b["T"] and a["R"] == b["R"]
    assert np.array_equal(a["P"], b["P"])                      # bit-exact likelihoods and compatibility sets
    assert np.array_equal(a["counts"], b["counts"])
    assert np.array_equal(a["new_to_old"], b["new_to_old"])
    assert np.array_equal(a["efflen"], b["efflen"])
    assert np.array_equal(a["complexity"], b["complexity"])


def _product(lib, loc, nce=False, mean=200.0, sd=20.0):
    model = lib.SequencingModel(mean, sd, int(loc.cand_length.max()))
    g = lib.build_collapsed_map(loc, model, no_coverage_exclude=nce)
    if g["return_code"] == 0:
        g["P"] = _dense(g)
    return g


def _gold_case(gold, ci):
    p = "case%d_" % ci
    loc = f1_py.RawLocus(*[gold[p + f] for f in make_golden_f1.FIELDS[:6]], int(gold[p + "strand_plus"]),
                         *[gold[p + f] for f in make_golden_f1.FIELDS[6:]])
    rc, T, R, nce = [int(x) for x in gold[p + "rc"]]
    want = dict(return_code=rc, T=T, R=R, P=gold[p + "P"], counts=gold[p + "counts"], new_to_old=gold[p + "new_to_old"],
                efflen=gold[p + "efflen"], complexity=gold[p + "complexity"])
    return loc, bool(nce), want, {int(k): int(v) for k, v in gold[p + "hist"]}


N_GOLD = len(make_golden_f1.CASES)


@pytest.mark.parametrize("ci", range(N_GOLD))
def test_golden_oracle_and_product(lib, orc, gold, ci):
    loc, nce, want, hist = _gold_case(gold, ci)
    regenerated, _ = make_golden_f1.make_case(make_golden_f1.CASES[ci])  # the committed inputs are what the generator makes
    assert np.array_equal(regenerated.read_pos, loc.read_pos) and np.array_equal(regenerated.cigar_len, loc.cigar_len)
    _same(orc.collapsed_map(loc, no_coverage_exclude=nce), want)
    _same(_product(lib, loc, nce), want)
    for lengths in (orc.fragment_lengths(loc), lib.fetch_fragment_lengths(loc)):
        assert dict(zip(*[x.tolist() for x in np.unique(lengths, return_counts=True)])) == hist


def test_golden_tables_and_model(lib, orc, gold):
    t = orc.sequencing_tables(200.0, 20.0, 4000)
    m = lib.SequencingModel(200.0, 20.0, 4000)
    for k in ("three_prime", "length_prob", "efflen"):
        assert np.array_equal(t[k], gold["tables_" + k])
    for l in (1, 2, 76, 199, 200, 201, 777, 4000):
        assert m.effective_length(l) == gold["tables_efflen"][l - 1]
        assert m.three_prime_prob(l, 4000) == gold["tables_three_prime"][l - 1]
        assert m.length_prob(min(l, 200), l) == gold["tables_length_prob"][l - 1]
    hist = {int(k): int(v) for k, v in gold["model_hist"]}
    assert orc.estimate_fragment_model(hist) == tuple(gold["model_est"])
    assert lib.estimate_fragment_model(hist) == tuple(gold["model_est"])
    with pytest.raises(lib.BayesError, match="1000"):
        lib.estimate_fragment_model({200: 10, 210: 5})


@pytest.mark.parametrize("seed", range(40))
def test_live_reference(lib, orc, ref, seed):
    rng = np.random.default_rng(seed)
    loc = f1_py.random_raw_locus(rng, n_exons=int(rng.integers(2, 10)), n_cand=int(rng.integers(1, 10)), n_frag=int(rng.integers(5, 700)),
                                 strand_plus=bool(seed % 2), two_subgraphs=seed % 3 == 0, indel_rate=0.05 if seed % 4 else 0.0)
    nce = seed % 5 == 0
    want = ref.collapsed_map(loc, no_coverage_exclude=nce)
    _same(orc.collapsed_map(loc, no_coverage_exclude=nce), want)
    _same(_product(lib, loc, nce), want)
    h = ref.fragment_length_histogram(loc)
    assert dict(zip(*[x.tolist() for x in np.unique(lib.fetch_fragment_lengths(loc), return_counts=True)])) == h


@pytest.mark.parametrize("seed,n_cand,n_frag", [(100, 2, 6000), (101, 3, 5000), (102, 12, 4000), (103, 1, 3000)])
def test_hashed_row_lookup_equals_linear_search(lib, orc, seed, n_cand, n_frag):
    # many fragments on few candidates: thousands of rows with the SAME support, the case the hash buckets must get right
    rng = np.random.default_rng(seed)
    loc = f1_py.random_raw_locus(rng, n_exons=5, n_cand=n_cand, n_frag=n_frag, premrna=seed % 2 == 0)
    _same(_product(lib, loc), orc.collapsed_map(loc))


def test_edge_cases_and_errors(lib, orc):
    rng = np.random.default_rng(1)
    # no fragment at all: every candidate is excluded for lack of mapped fragments -> return code 3 (assembler.cpp:1521)
    loc = f1_py.random_raw_locus(rng, n_cand=3, n_frag=0)
    assert _product(lib, loc)["return_code"] == 3 and orc.collapsed_map(loc)["return_code"] == 3
    # only junk fragments with coverage exclusion off -> no rows -> return code 4
    loc = f1_py.random_raw_locus(rng, n_cand=2, n_frag=30, junk_rate=1.0, premrna=False)
    loc.read_pos[:] += 100000
    assert _product(lib, loc, nce=True)["return_code"] == 4 and orc.collapsed_map(loc, no_coverage_exclude=True)["return_code"] == 4
    # malformed input is refused with a message (the reference asserts / exits, alignmentParser.cpp:67-68, 158-162)
    loc = f1_py.random_raw_locus(rng, n_cand=2, n_frag=10)
    model = lib.SequencingModel(200.0, 20.0, int(loc.cand_length.max()))
    bad = f1_py.random_raw_locus(np.random.default_rng(2), n_cand=2, n_frag=10)
    bad.cigar_op[0] = ord("S")
    with pytest.raises(lib.BayesError, match="CIGAR"):
        lib.build_collapsed_map(bad, model)
    short = lib.SequencingModel(200.0, 20.0, 50)
    with pytest.raises(lib.BayesError, match="longer"):
        lib.build_collapsed_map(loc, short)


def test_collapsed_map_feeds_the_sampler_interface(lib, oracle, orc):
    # the CSR arrays of the builder are a valid bayes_locus_desc: the host pieces of the sampler accept them
    rng = np.random.default_rng(4)
    raw = f1_py.random_raw_locus(rng, n_exons=7, n_cand=6, n_frag=800)
    g = _product(lib, raw)
    assert g["return_code"] == 0
    loc = lib.Locus(g["T"], g["row_ptr"], g["col_idx"], g["val"], g["counts"], g["efflen"], 1e6)
    from oracle.oracle_py import SITE
    n, flags = lib.min_read_cover(loc, 99)
    assert n == oracle.min_read_cover(loc, SITE, 99)[0] and 1 <= n <= g["T"]
    assert g["stats"][0] == raw.n_frag and g["stats"][5] == g["R"] and g["stats"][1] + g["stats"][2] == raw.n_frag
    assert g["counts"].sum() <= g["stats"][1]
