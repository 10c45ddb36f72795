"""Pin the CPU oracle (oracle/bayes_oracle.c) before anything is compared with it.

 * against tests/golden/ref_vectors.npz -- outputs of the UNMODIFIED reference translation units (oracle/_ref) recorded by
   tests/golden/make_golden.py in the build container (the reference has no golden vectors of its own);
 * against the reference itself, live, where oracle/_ref exists (skipped on a box with neither /root/reference nor a
   prebuilt libbayesref.so);
 * known-answer tests of the two word generators (MT19937: Matsumoto & Nishimura's 10000th output; Philox4x32-10:
   Random123 kat_vectors) and closed forms of the domain (T == 1, simplex table vs scipy).

Integer results and RNG word counts must be identical; floating point agrees to 1e-12 relative.
"""
import os

import numpy as np
import pytest

from oracle.oracle_py import SEQ, SITE, Locus, random_locus

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz")
RTOL = 1e-12


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _case(gold, ci):
    T, R, gamma, n_total, seed, burn, samples = gold["meta"][ci]
    p = "case%d_" % ci
    loc = Locus(int(T), gold[p + "row_ptr"], gold[p + "col_idx"], gold[p + "val"], gold[p + "row_count"], gold[p + "efflen"],
                n_total, gamma)
    return p, loc, int(seed), int(burn), int(samples)


N_CASES = 9


# ---------------------------------------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("ci", range(N_CASES))
def test_golden_cover_and_init(oracle, gold, ci):
    p, loc, seed, _, _ = _case(gold, ci)
    n, _, w = oracle.min_read_cover(loc, SEQ, seed)
    assert [n, w] == gold[p + "cover"].tolist()
    c, w = oracle.init_assignment(loc, SEQ, seed)
    assert np.array_equal(c, gold[p + "init_counts"]) and w == int(gold[p + "init_words"])
    assert c.sum() == loc.n_frag


@pytest.mark.parametrize("ci", range(1, N_CASES))
def test_golden_assignment_and_expression(oracle, gold, ci):
    p, loc, seed, _, _ = _case(gold, ci)
    c, w = oracle.generate_assignment(loc, gold[p + "theta_in"], SEQ, seed)
    assert np.array_equal(c, gold[p + "assign_counts"]) and w == int(gold[p + "assign_words"])
    th, order, w = oracle.generate_expression(c, int(gold[p + "expr_b"]), loc.gamma, SEQ, seed)
    assert np.array_equal(order, gold[p + "expr_order"]) and w == int(gold[p + "expr_words"])
    np.testing.assert_allclose(th, gold[p + "expr_theta"], rtol=RTOL)
    assert (th > 0).sum() == int(gold[p + "expr_b"])


@pytest.mark.parametrize("ci", range(N_CASES))
def test_golden_full_sampler(oracle, gold, ci):
    p, loc, seed, burn, samples = _case(gold, ci)
    o = oracle.run_sampler(loc, burn, samples, SEQ, seed)
    assert np.array_equal(o["occ"], gold[p + "occ"])
    np.testing.assert_allclose(o["mean"], gold[p + "mean"], rtol=RTOL)
    np.testing.assert_allclose(o["m2"], gold[p + "m2"], rtol=1e-9, atol=1e-300)
    assert o["pi"] == float(gold[p + "pi"])
    assert o["words"] == int(gold[p + "run"][1])


def test_golden_simplex_tables(oracle, gold):
    for ti, (pi, gamma, T, nf) in enumerate(gold["tables"]):
        tab, rl = oracle.simplex_table(pi, gamma, int(T), int(nf))
        assert np.array_equal(rl, gold["table%d_len" % ti])
        np.testing.assert_allclose(tab, gold["table%d_val" % ti], rtol=RTOL, atol=1e-300)


def test_golden_top_marginals(oracle, gold):
    idx, mo, sd = oracle.top_marginals(100, gold["top_occ"], gold["top_mean"], gold["top_m2"], 0.5)
    assert np.array_equal(idx, gold["top_idx"])
    np.testing.assert_allclose(mo, gold["top_mean_out"], rtol=RTOL)
    np.testing.assert_allclose(sd, gold["top_sd_out"], rtol=RTOL)


# ---------------------------------------------------------------------------------------------- live reference
LIVE = [(2, 5, 30, 1.0), (3, 40, 60, 1.0), (7, 70, 19, 1.0), (10, 150, 300, 0.5), (31, 200, 100, 1.0), (33, 100, 2000, 2.0),
        (64, 150, 60, 1.0)]


@pytest.mark.parametrize("T,R,mc,gamma", LIVE)
def test_live_steps(oracle, reference, T, R, mc, gamma):
    rng = np.random.default_rng(T * 131 + R)
    loc = random_locus(rng, T, R, max_count=mc, gamma=gamma)
    for seed in (1, 12345, 2 ** 31 - 5):
        n, _, w = oracle.min_read_cover(loc, SEQ, seed)
        assert (n, w) == reference.min_read_cover(loc, seed)
        c, w = oracle.init_assignment(loc, SEQ, seed)
        rc, rw = reference.init_assignment(loc, seed)
        assert np.array_equal(c, rc) and w == rw
        th = rng.dirichlet(np.ones(T))
        th[rng.random(T) < 0.3] = 0.0
        P = loc.dense()
        for i in np.nonzero((P * th).sum(1) <= 0)[0]:
            th[loc.col_idx[loc.row_ptr[i]]] = 0.05
        th /= th.sum()
        c, w = oracle.generate_assignment(loc, th, SEQ, seed)
        rc, rw = reference.generate_assignment(loc, th, seed)
        assert np.array_equal(c, rc) and w == rw
        n_plus = int((c > 0).sum())
        for b in sorted({n_plus, min(T, n_plus + 2), T}):
            e, order, w = oracle.generate_expression(c, b, gamma, SEQ, seed)
            re_, rorder, rw = reference.generate_expression(c, loc.n_frag, b, gamma, seed)
            assert np.array_equal(order, rorder) and w == rw
            np.testing.assert_allclose(e, re_, rtol=RTOL)


@pytest.mark.parametrize("T,R,mc,gamma", LIVE)
def test_live_full_sampler(oracle, reference, T, R, mc, gamma):
    rng = np.random.default_rng(T * 17 + R)
    loc = random_locus(rng, T, R, max_count=mc, gamma=gamma)
    seed = int(rng.integers(1, 2 ** 31 - 1))
    burn, samples = 40 + 2 * T, 400 + 20 * T
    o = oracle.run_sampler(loc, burn, samples, SEQ, seed)
    r = reference.run_locus(loc, burn, samples, seed)
    assert np.array_equal(o["occ"], r["occ"])
    np.testing.assert_allclose(o["mean"], r["mean"], rtol=RTOL)
    np.testing.assert_allclose(o["m2"], r["m2"], rtol=1e-9, atol=1e-300)
    assert o["pi"] == r["pi"] and o["words"] == r["words"]


def test_live_simplex(oracle, reference):
    for (pi, gamma, T, nf) in [(0.3, 1.0, 6, 50), (0.9, 0.7, 17, 1234), (0.02, 1.0, 80, 20000)]:
        tab, rl = oracle.simplex_table(pi, gamma, T, nf)
        rtab, rrl = reference.simplex_table(pi, gamma, T, nf)
        assert np.array_equal(rl, rrl)
        np.testing.assert_allclose(tab, rtab, rtol=RTOL, atol=1e-300)
        for k in (1, max(1, T // 2), T):
            for seed in (3, 99):
                assert oracle.sample_simplex(tab, rl, k, SEQ, seed) == reference.sample_simplex(pi, gamma, T, nf, k, seed)


# ---------------------------------------------------------------------------------------------- known answers
def test_mt19937_known_answer(oracle):
    # Matsumoto & Nishimura / ISO C++ [rand.predef]: the 10000th output of mt19937 seeded with 5489 is 4123659995
    out = oracle.mt(5489, 10000)
    assert int(out[-1]) == 4123659995
    assert int(out[0]) == 3499211612


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds: ctr[4], key[2] -> out[4]
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        got = oracle.philox(key[0] | (key[1] << 32), ctr)
        assert [int(x) for x in got] == want


def test_single_candidate_closed_form(oracle):
    # gibbsSampler.cpp:57-67: `samples` identical samples of N_f 1e9 / (N_total efflen)
    loc = Locus(1, [0, 1, 2], [0, 0], [1.0, 1.0], [7, 5], [1500.0], 1000000)
    o = oracle.run_sampler(loc, 10, 100, SEQ, 1)
    assert o["occ"][0] == 100 and o["m2"][0] == 0
    assert o["mean"][0] == (12 * 1e9) / (1000000 * 1500.0)


def test_simplex_table_against_scipy(oracle):
    from scipy.special import gammaln, logsumexp
    pi, gamma, T, nf = 0.35, 1.0, 12, 300
    tab, rl = oracle.simplex_table(pi, gamma, T, nf)
    for k in range(1, T + 1):
        j = np.arange(k, T + 1)
        lp = (gammaln(T - k + 1) - gammaln(j - k + 1) - gammaln(T - j + 1) + j * np.log(pi) + (T - j) * np.log(1 - pi)
              + gammaln(j * gamma) - gammaln(nf + j * gamma))
        cdf = np.exp(np.logaddexp.accumulate(lp) - logsumexp(lp))
        n = rl[k - 1]
        np.testing.assert_allclose(tab[k - 1, :n], cdf[:n], rtol=1e-9)
        assert tab[k - 1, n - 1] > 1 - 1e-12


def test_sampler_invariants(oracle):
    # counts sum to N_f after every assignment, S+_counts within S+_expr, |S+_expr| == b, theta sums to 1
    rng = np.random.default_rng(5)
    loc = random_locus(rng, 9, 80, max_count=100)
    o = oracle.run_sampler(loc, 30, 300, SITE, 0xABCDEF, trace_iters=330)
    assert np.all(o["counts_trace"].sum(1) == loc.n_frag)
    assert np.allclose(o["theta_trace"].sum(1), 1.0, atol=1e-12)
    assert np.array_equal((o["theta_trace"] > 0).sum(1), o["b_trace"])
    prev = np.vstack([o["init_counts"][None, :], o["counts_trace"][:-1]])
    assert np.all((prev > 0) <= (o["theta_trace"] > 0))
    assert np.all(o["occ"] == (o["theta_trace"][30:] > 0).sum(0))
