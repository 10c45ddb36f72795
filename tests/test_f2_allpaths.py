"""Candidate enumeration: bayes_all_paths_build against
 (a) tests/golden/f2_vectors.json = answers of the UNMODIFIED reference allPaths.cpp (tests/golden/make_golden_f2.py),
 (b) the live oracle/_ref/libbayesref_f2.so where it exists,
 (c) the plain-Python restatement oracle/f2_py.all_paths_oracle (which is itself pinned against (a)).
Indices, coordinates, names and log numbers are compared exactly (north_star: path enumeration bit-exact)."""
import json
import os

import numpy as np
import pytest

from oracle import f2_py

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "f2_vectors.json")))
ARRAYS = ("exon_ptr", "exon_start", "exon_end", "cand_length", "is_premrna")


def _case(c):
    if "gen" in c:
        seed, n_exons, p_edge, pre, nopre, maxc = c["gen"]
        return f2_py.random_splice_graph(seed, n_exons, p_edge=p_edge, pre_mrna=pre), nopre, maxc
    return c["dot"], c["no_pre_mrna"], c["max_candidate_number"]


def _same(got, want, what=""):
    assert got["names"] == list(want["names"]), what
    for k in ARRAYS:
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), (what, k)
    assert [int(x) for x in np.asarray(got["stats"])[:5]] == [int(x) for x in np.asarray(want["stats"])[:5]], what
    if got["names"]:
        assert got["graph_name"] == want["graph_name"]


@pytest.fixture(scope="module")
def ref_f2():
    from oracle import oracle_py
    oracle_py.build(ref=True)
    if not os.path.exists(f2_py.REF_F2_SO):
        pytest.skip("oracle/_ref/libbayesref_f2.so not built (no /root/reference here)")
    return f2_py.ReferenceF2()


@pytest.mark.parametrize("ci", range(len(GOLDEN)))
def test_oracle_restatement_pinned_against_golden(ci):
    dot, nopre, maxc = _case(GOLDEN[ci])
    _same(f2_py.flatten(f2_py.all_paths_oracle(dot, nopre, maxc)), GOLDEN[ci]["want"], ci)


@pytest.mark.parametrize("ci", range(len(GOLDEN)))
def test_product_against_golden(lib, ci):
    dot, nopre, maxc = _case(GOLDEN[ci])
    got = lib.all_paths(dot, nopre, maxc)
    _same(got, GOLDEN[ci]["want"], ci)
    assert got["stats"][5] == got["stats"][3] + 1  # final junction threshold = 1 + adjustments (allPaths.cpp:47, :213)


def test_golden_matches_live_reference(ref_f2):
    for ci, c in enumerate(GOLDEN):
        dot, nopre, maxc = _case(c)
        _same(ref_f2.all_paths(dot, nopre, maxc), c["want"], ci)


@pytest.mark.parametrize("block", range(6))
def test_product_against_live_reference_random(lib, ref_f2, block):
    for seed in range(50 * block, 50 * block + 50):
        n_exons = 2 + seed % 19
        dot = f2_py.random_splice_graph(seed, n_exons, p_edge=0.25 + 0.05 * ((seed * 7) % 13), pre_mrna=seed % 3 != 0,
                                        strand="+" if seed % 2 else "-")
        for maxc in (100, 7, 1000):
            for nopre in (False, True):
                _same(lib.all_paths(dot, nopre, maxc), ref_f2.all_paths(dot, nopre, maxc), (seed, maxc, nopre))


def test_product_against_restatement_random(lib):
    """Runs everywhere (no reference needed): small graphs, every pruning regime."""
    for seed in range(500, 620):
        dot = f2_py.random_splice_graph(seed, 2 + seed % 10, p_edge=0.3 + 0.05 * (seed % 11), pre_mrna=seed % 4 != 0)
        for maxc in (100, 4, 1):
            _same(lib.all_paths(dot, seed % 2 == 0, maxc), f2_py.flatten(f2_py.all_paths_oracle(dot, seed % 2 == 0, maxc)), (seed, maxc))


def test_vertex_order_is_by_id_string(lib):
    """Twelve isolated exons: candidates come out in the order "0" "1" "10" "11" "2" ... (read_graphviz's node map)."""
    exons = [(100 * i + 1, 100 * i + 50) for i in range(12)]
    got = lib.all_paths(f2_py.dot_text("g", "chr1", "+", exons, [], pre_mrna=False), False, 100)
    order = sorted(range(12), key=str)
    assert list(got["exon_start"]) == [exons[i][0] for i in order]
    assert got["names"] == ["g_seq%d" % i for i in range(12)]


def test_config3_class_graph(lib):
    """A graph with thousands of paths pruned to --max-candidate-number 1000 (BASELINE.json configs[2]): candidate count in
    range, every candidate a strictly increasing exon chain, lengths consistent, the pre-mRNA last (its id sorts last)."""
    dot = f2_py.random_splice_graph(4242, 45, p_edge=0.7, max_cov=300)
    got = lib.all_paths(dot, False, 1000)
    n = got["n_cand"]
    assert 1 < n <= 1000 and got["stats"][3] > 0
    for i in range(n):
        s = got["exon_start"][got["exon_ptr"][i]:got["exon_ptr"][i + 1]]
        e = got["exon_end"][got["exon_ptr"][i]:got["exon_ptr"][i + 1]]
        assert np.all(e >= s) and np.all(s[1:] > e[:-1] + 1)
        assert got["cand_length"][i] == int((e - s + 1).sum())
    assert got["is_premrna"][-1] == 1 and got["is_premrna"][:-1].sum() == 0 and got["names"][-1].endswith("_pre")
    # one more allowed candidate than the unpruned count changes nothing; the threshold found is minimal
    full = lib.all_paths(dot, False, 10 ** 9)
    assert full["stats"][3] == 0 and full["n_cand"] > 1000
    relaxed = lib.all_paths(dot, False, n)
    assert relaxed["names"] == got["names"] and relaxed["stats"][5] == got["stats"][5]


def test_paths_feed_the_likelihood_builder(lib):
    """f2 -> f1: the exported arrays are bayes_raw_locus' candidate arrays as they are."""
    from oracle import f1_py
    dot = f2_py.random_splice_graph(77, 7, p_edge=0.6)
    ps = lib.all_paths(dot, False, 100)
    rng = np.random.default_rng(3)
    raw = f1_py.fragments_for_candidates(rng, ps["exon_ptr"], ps["exon_start"], ps["exon_end"], ps["cand_length"], ps["is_premrna"],
                                         n_frag=300, strand_plus=ps["strand"] == "+")
    model = lib.SequencingModel(200.0, 20.0, int(ps["cand_length"].max()))
    cm = lib.build_collapsed_map(raw, model)
    assert cm["return_code"] == 0 and cm["T"] >= 1 and cm["counts"].sum() > 0
    want = f1_py.OracleF1().collapsed_map(raw, max_len=model.max_len)
    assert np.array_equal(cm["counts"], want["counts"]) and np.array_equal(cm["new_to_old"], want["new_to_old"])


@pytest.mark.parametrize("dot,msg", [
    ("graph g {\n}\n", "digraph"),
    ("digraph g {\n0 [label=\"c;+;1;10\"]\n", "ends before"),
    ("digraph g {\n0 [label=\"c;+;1;10\"]\n1 [label=\"d;+;20;30\"]\n}\n", "different references"),
    ("digraph g {\n0 [label=\"c;+;1;10\"]\n0->1 [coverage=\"3\"]\n}\n", "no label"),
    ("digraph g {\n0 [label=\"c;+;10;1\"]\n}\n", "ends before it starts"),
    ("digraph g {\n0 [label=\"c;+;1;10\"]\n1 [label=\"c;+;20;30\"]\n0->1 [coverage=\"3\"]\n1->0 [coverage=\"3\"]\n}\n", "cycle"),
])
def test_malformed_graphs_are_errors_not_aborts(lib, dot, msg):
    """The reference asserts (allPaths.cpp:73, :79, :110-118); the library returns BAYES_E_INVALID with a message."""
    with pytest.raises(lib.BayesError) as e:
        lib.all_paths(dot, False, 100)
    assert msg in str(e.value)
