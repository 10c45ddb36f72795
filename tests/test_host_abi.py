"""CPU-side tests of the product library: the C ABI loads and exports every symbol include/*.h declares, the host pieces
of the path (minimum read cover, simplex CDF table, chain keys, cost model, LPT partition, top marginals, Philox) agree
with the oracle, and there is NO CPU fallback (creating a context without a GPU fails with BAYES_E_NODEVICE).

No device compute is called here; the parity tests proper are tests/test_gpu_parity.py (-m gpu).
"""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from oracle.oracle_py import SITE, random_locus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(bayes_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(lib):
    L = lib.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(lib.EXPORTED_SYMBOLS) == declared  # the ctypes binding covers the whole header


def test_desc_layout_matches_header(lib):
    # bayes_locus_desc: 2 x i32, 5 pointers, 3 doubles, 2 x i32, u64, 2 x i32
    assert C.sizeof(lib.LocusDesc) == 8 + 5 * 8 + 3 * 8 + 8 + 8 + 8
    from bayesembler_b200 import synth
    assert synth.DESC_DTYPE.itemsize == C.sizeof(lib.LocusDesc)
    for name, _ in lib.LocusDesc._fields_:
        assert synth.DESC_DTYPE.fields[name][1] == getattr(lib.LocusDesc, name).offset


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device error path cannot be exercised")
    L = lib.load_library()
    h = C.c_void_p()
    rc = L.bayes_gibbs_create(0, C.byref(h))
    assert rc == -5 and not h.value  # BAYES_E_NODEVICE
    assert b"no CPU path" in L.bayes_last_error()
    with pytest.raises(lib.BayesError):
        lib.GibbsContext(0)
    rng = np.random.default_rng(0)
    with pytest.raises(lib.BayesError):
        lib.step_init_assignment(random_locus(rng, 3, 5), 1)


def test_missing_extension_fails_loudly(lib, tmp_path):
    with pytest.raises(lib.BayesError):
        lib.load_library(str(tmp_path / "libbayesembler_b200.so"))


@pytest.mark.parametrize("T,R", [(1, 3), (2, 5), (5, 40), (12, 200), (33, 150), (100, 400), (300, 500)])
def test_min_read_cover_matches_oracle(lib, oracle, T, R):
    rng = np.random.default_rng(T * 7 + R)
    loc = random_locus(rng, T, R, density=0.15)
    for key in (1, 0xDEADBEEFCAFE, 2 ** 63 + 11):
        n, flags = lib.min_read_cover(loc, key)
        on, oflags, _ = oracle.min_read_cover(loc, SITE, key)
        assert n == on and np.array_equal(flags, oflags)
        P = loc.dense()
        assert np.all((P[:, flags > 0] >= np.finfo(float).tiny).any(1))  # it is a cover


@pytest.mark.parametrize("pi,gamma,T,nf", [(0.5, 1.0, 1, 10), (0.5, 1.0, 2, 10), (0.25, 1.0, 8, 500), (1 - 2.0 ** -52, 1.0, 5, 100),
                                            (0.1, 0.5, 30, 2000), (0.03, 2.0, 100, 9000), (0.01, 1.0, 700, 12000)])
def test_simplex_table_matches_oracle(lib, oracle, pi, gamma, T, nf):
    row_off, tab = lib.simplex_table(pi, gamma, T, nf)
    otab, orl = oracle.simplex_table(pi, gamma, T, nf)
    assert np.array_equal(np.diff(row_off), orl)
    for k in range(T):
        # bit-identical: the O(T) lgamma look-ups are the same lgamma values the reference evaluates in place
        assert np.array_equal(tab[row_off[k]:row_off[k + 1]], otab[k, :orl[k]])


def test_golden_simplex_tables_through_abi(lib):
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
    for ti, (pi, gamma, T, nf) in enumerate(g["tables"]):
        row_off, tab = lib.simplex_table(pi, gamma, int(T), int(nf))
        assert np.array_equal(np.diff(row_off), g["table%d_len" % ti])
        for k in range(int(T)):
            np.testing.assert_allclose(tab[row_off[k]:row_off[k + 1]], g["table%d_val" % ti][k, :row_off[k + 1] - row_off[k]],
                                       rtol=1e-12, atol=1e-300)


def test_top_marginals_golden(lib, oracle):
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
    idx, mo, sd = lib.top_marginals(100, g["top_occ"], g["top_mean"], g["top_m2"], 0.5)
    assert np.array_equal(idx, g["top_idx"])
    np.testing.assert_allclose(mo, g["top_mean_out"], rtol=1e-15)
    np.testing.assert_allclose(sd, g["top_sd_out"], rtol=1e-15)


def test_philox_known_answers(lib, oracle):
    kat = [([0, 0, 0, 0], 0, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, 0xffffffffffffffff, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], 0xa4093822 | (0x299f31d0 << 32), [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert [int(x) for x in lib.philox4x32(key, ctr)] == want
    rng = np.random.default_rng(1)
    for _ in range(50):
        key = int(rng.integers(0, 2 ** 63))
        ctr = rng.integers(0, 2 ** 32, 4).astype(np.uint32)
        assert np.array_equal(lib.philox4x32(key, ctr), oracle.philox(key, ctr))


def test_chain_keys(lib):
    assert lib.chain_key(12345, 0) == 12345  # chain 0 = the reference's one chain, keyed by graph_seed itself
    keys = {lib.chain_key(12345, c) for c in range(4096)}
    assert len(keys) == 4096
    assert lib.chain_key(12345, 1) != lib.chain_key(12346, 1)


def test_cost_model_and_lpt(lib):
    rng = np.random.default_rng(3)
    loc = random_locus(rng, 10, 100)
    c1 = lib.locus_cost(loc, 100, 1000)
    assert c1 == 1100 * (loc.nnz + 4 * loc.R + 8 * loc.T)
    assert lib.locus_cost(loc, 100, 1000, n_chains=8) == 8 * c1
    cost = rng.pareto(1.3, 5000) + 0.01
    for parts in (1, 2, 3, 8):
        p = lib.partition_lpt(cost, parts)
        assert p.min() >= 0 and p.max() < parts
        load = np.bincount(p, weights=cost, minlength=parts)
        # LPT guarantee: makespan <= (4/3 - 1/(3m)) OPT, and OPT >= max(mean load, largest item)
        assert load.max() <= (4.0 / 3.0) * max(cost.sum() / parts, cost.max()) + 1e-9
    assert np.array_equal(lib.partition_lpt(cost, 4), lib.partition_lpt(cost, 4))  # deterministic
    assert lib.partition_lpt(np.array([3.0, 1.0, 2.0]), 2).tolist() == [0, 1, 1]


def test_synth_batch_is_regenerable(lib):
    from bayesembler_b200 import synth
    a = synth.SynthBatch(2, n_loci=64, n_threads=2)
    b = synth.SynthBatch(2, ids=np.arange(20, 40), n_threads=3)
    for j, i in enumerate(range(20, 40)):
        la, lb = a.locus(i), b.locus(j)
        assert la.T == lb.T and np.array_equal(la.row_ptr, lb.row_ptr) and np.array_equal(la.col_idx, lb.col_idx)
        assert np.array_equal(la.val, lb.val) and np.array_equal(la.row_count, lb.row_count)
    # CollapsedMap invariants (alignmentParser.cpp:369, 373-411): rows normalised, counts >= 1, columns ascending
    for i in range(a.n):
        l = a.locus(i)
        assert np.all(l.row_count >= 1) and l.row_ptr[0] == 0
        for r in range(l.R):
            s, e = l.row_ptr[r], l.row_ptr[r + 1]
            assert e > s and np.all(np.diff(l.col_idx[s:e]) > 0) and abs(l.val[s:e].sum() - 1) < 1e-12
    st = a.stats()
    assert st["draws"] > 0 and st["algo_bytes"] > 0


def _bad_locus(lib, mutate):
    rng = np.random.default_rng(9)
    loc = random_locus(rng, 4, 6, single_frac=0.0)
    mutate(loc)
    return loc


@pytest.mark.parametrize("name,mutate", [
    ("columns not ascending", lambda l: l.col_idx.__setitem__(slice(0, 2), l.col_idx[:2][::-1].copy())),
    ("column out of range", lambda l: l.col_idx.__setitem__(1, 99)),
    ("negative probability", lambda l: l.val.__setitem__(0, -0.1)),
    ("nan probability", lambda l: l.val.__setitem__(0, float("nan"))),
    ("row count < 1", lambda l: l.row_count.__setitem__(2, 0)),
    ("row without support", lambda l: l.val.__setitem__(slice(int(l.row_ptr[1]), int(l.row_ptr[2])), 0.0)),
    ("row_ptr not monotone", lambda l: l.row_ptr.__setitem__(2, 0)),
])
def test_malformed_loci_are_rejected(lib, name, mutate):
    # the reference aborts on such input through asserts (assignmentGenerator.cpp:67, 284); the library returns
    # BAYES_E_INVALID and a message instead
    loc = _bad_locus(lib, mutate)
    d = lib.make_desc(loc, 0, 1, 0)
    L = lib.load_library()
    rc = L.bayes_min_read_cover(C.byref(d), 1, None)
    assert rc == -1, name
    assert len(L.bayes_last_error()) > 0


def test_degenerate_but_valid_loci(lib, oracle):
    # one row / one candidate / every row with a single cell / wide row: the host pieces accept them
    from oracle.oracle_py import Locus
    one = Locus(1, [0, 1], [0], [1.0], [5], [1000.0], 1000)
    assert lib.min_read_cover(one, 7)[0] == 1
    single = Locus(3, [0, 1, 2, 3], [2, 0, 1], [1.0, 1.0, 1.0], [4, 1, 9], [500.0, 600.0, 700.0], 1000)
    n, flags = lib.min_read_cover(single, 7)
    assert n == 3 and flags.tolist() == [1, 1, 1]
    assert lib.min_read_cover(single, 7)[0] == oracle.min_read_cover(single, SITE, 7)[0]
    T = 4096  # the supported maximum (internal.h kMaxT)
    wide = Locus(T, [0, T], np.arange(T), np.full(T, 1.0 / T), [100], np.full(T, 900.0), 1000)
    assert lib.min_read_cover(wide, 3)[0] == 1
    row_off, tab = lib.simplex_table(1.0 / T, 1.0, T, 100)
    assert len(row_off) == T + 1 and np.all(np.diff(row_off) >= 1) and tab[row_off[1] - 1] > 1 - 1e-9
