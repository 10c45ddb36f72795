"""The C++ host facade (bayesembler_b200/host/bayesembler_b200.hpp) -- the reference's class surface on top of the C ABI.

CPU: the GTF writers and getTopMarginals against text recorded from the reference's own GibbsOutput
(tests/golden/ref_vectors.npz, gibbsOutput.cpp:77-166) and against the live reference where oracle/_ref exists; the
facade refuses to sample without a GPU.  GPU: a worker written like assembler.cpp:1548-1609 on the facade classes gives
the same accumulators as the C ABI called directly, in both the per-locus and the batched form.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py
from oracle.oracle_py import random_locus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "facade_driver.cpp")
HPP = os.path.join(ROOT, "bayesembler_b200", "host", "bayesembler_b200.hpp")
OUT = os.path.join(ROOT, "tests", "cpp", "libfacade_test.so")
GOLD = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
GTF_SHAPES = [dict(T=7, samples=40, n_cand=9), dict(T=3, samples=25, n_cand=3), dict(T=12, samples=60, n_cand=20),
              dict(T=1, samples=10, n_cand=2)]


@pytest.mark.parametrize("gi", range(4))
def test_gtf_text_matches_reference_golden(facade, gi):
    g = np.load(GOLD)
    case = oracle_py.gtf_case(100 + gi, **GTF_SHAPES[gi])
    assert oracle_py.call_gtf(facade.facade_gtf, case, 0) == bytes(g["gtf%d_ensemble" % gi]).decode()
    assert oracle_py.call_gtf(facade.facade_gtf, case, 1) == bytes(g["gtf%d_candidates" % gi]).decode()


def test_gtf_text_matches_live_reference(facade, reference):
    for seed in range(20):
        case = oracle_py.gtf_case(seed, T=1 + seed % 9, samples=10 + 7 * seed, n_cand=10)
        case["count_threshold"] = [0.0, 12.0, 500.0][seed % 3]
        for mode in (0, 1):
            assert oracle_py.call_gtf(facade.facade_gtf, case, mode) == reference.gtf(case, mode)


def test_facade_simplex_table(facade, oracle):
    pi, gamma, T, nf = 0.2, 1.0, 25, 800
    rl = np.zeros(T, np.int32)
    flat = np.zeros(T * T)
    facade.facade_simplex_rows.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = facade.facade_simplex_rows(pi, gamma, T, nf, rl.ctypes.data, flat.ctypes.data)
    otab, orl = oracle.simplex_table(pi, gamma, T, nf)
    assert np.array_equal(rl, orl) and n == orl.sum()
    at = 0
    for k in range(T):
        assert np.array_equal(flat[at:at + rl[k]], otab[k, :rl[k]])
        at += rl[k]


def _run_loci(facade, loci, seeds, burn, samples, batched):
    n = len(loci)
    T_arr = np.array([l.T for l in loci], np.int32)
    row_off = np.zeros(n + 1, np.int64)
    cell_off = np.zeros(n + 1, np.int64)
    t_off = np.zeros(n + 1, np.int64)
    for i, l in enumerate(loci):
        row_off[i + 1], cell_off[i + 1], t_off[i + 1] = row_off[i] + l.R, cell_off[i] + l.nnz, t_off[i] + l.T
    cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs), dt)  # noqa: E731
    row_ptr, col, val = cat([l.row_ptr for l in loci], np.int32), cat([l.col_idx for l in loci], np.int32), cat([l.val for l in loci], np.float64)
    cnt, eff = cat([l.row_count for l in loci], np.int32), cat([l.efflen for l in loci], np.float64)
    occ, mean, m2, pi = np.zeros(t_off[-1], np.int32), np.zeros(t_off[-1]), np.zeros(t_off[-1]), np.zeros(n)
    err = C.create_string_buffer(512)
    vp = C.c_void_p
    facade.facade_run_loci.restype = C.c_int
    facade.facade_run_loci.argtypes = [C.c_int] + [vp] * 9 + [C.c_double, C.c_double, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_char_p, C.c_int]
    p = lambda a: a.ctypes.data  # noqa: E731
    seeds_a, burn_a, samples_a = np.ascontiguousarray(seeds, np.uint64), np.ascontiguousarray(burn, np.int32), np.ascontiguousarray(samples, np.int32)
    rc = facade.facade_run_loci(n, p(T_arr), p(row_off), p(cell_off), p(t_off), p(row_ptr), p(col), p(val), p(cnt), p(eff),
                                loci[0].n_total, loci[0].gamma, p(seeds_a), p(burn_a), p(samples_a), int(batched),
                                p(occ), p(mean), p(m2), p(pi), err, len(err))
    return rc, err.value.decode(), [dict(occ=occ[t_off[i]:t_off[i + 1]], mean=mean[t_off[i]:t_off[i + 1]], m2=m2[t_off[i]:t_off[i + 1]],
                                         pi=pi[i]) for i in range(n)]


def test_facade_has_no_cpu_sampler(facade):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rng = np.random.default_rng(0)
    for batched in (0, 1):
        rc, err, _ = _run_loci(facade, [random_locus(rng, 3, 6)], [1], [5], [50], batched)
        assert rc == -1 and "no CPU path" in err


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [0, 1])
def test_facade_worker_matches_c_abi(facade, lib, batched):
    rng = np.random.default_rng(77)
    shapes = [(1, 3, 9), (2, 6, 40), (5, 50, 100), (12, 70, 30), (40, 120, 200)]
    loci = [random_locus(rng, T, R, max_count=mc) for (T, R, mc) in shapes]
    seeds = [int(s) for s in rng.integers(1, 2 ** 31, len(loci))]
    burn = [30 + 3 * l.T for l in loci]
    samples = [10 * b for b in burn]
    rc, err, got = _run_loci(facade, loci, seeds, burn, samples, batched)
    assert rc == 0, err
    ctx = lib.GibbsContext(0)
    ctx.run_batch(lib.make_desc_array(loci, burn, samples, seeds))
    for i in range(len(loci)):
        want = ctx.fetch(i)
        assert got[i]["pi"] == ctx.locus_info(i)["pi"]
        assert np.array_equal(got[i]["occ"], want["occ"]), i
        assert np.array_equal(got[i]["mean"], want["mean"]), (i, got[i]["mean"], want["mean"])
        assert np.array_equal(got[i]["m2"], want["m2"]), (i, got[i]["m2"], want["m2"])
    ctx.close()


@pytest.mark.parametrize("ci", range(9))
def test_facade_likelihood_builder_matches_reference_golden(facade, ci):
    """AlignmentParser / SequencingModel / GaussianFragmentLengthModel of the facade (row f1) against the reference's
    CollapsedMaps recorded in tests/golden/f1_vectors.npz."""
    import sys
    from oracle import f1_py
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_f1
    g = np.load(os.path.join(ROOT, "tests", "golden", "f1_vectors.npz"))
    p = "case%d_" % ci
    loc = f1_py.RawLocus(*[g[p + f] for f in make_golden_f1.FIELDS[:6]], int(g[p + "strand_plus"]),
                         *[g[p + f] for f in make_golden_f1.FIELDS[6:]])
    rc, T, R, nce = [int(x) for x in g[p + "rc"]]
    fn = facade.facade_collapsed_map
    fn.restype = C.c_int
    fn.argtypes = f1_py._CM_ARGS
    got = f1_py._call_collapsed_map(fn, loc, 200.0, 20.0, int(loc.cand_length.max()), bool(nce))
    assert got["return_code"] == rc
    if rc == 0:
        assert got["T"] == T and got["R"] == R
        assert np.array_equal(got["P"], g[p + "P"]) and np.array_equal(got["counts"], g[p + "counts"])
        assert np.array_equal(got["new_to_old"], g[p + "new_to_old"]) and np.array_equal(got["efflen"], g[p + "efflen"])
        assert np.array_equal(got["complexity"], g[p + "complexity"])


def test_facade_all_paths_matches_reference_golden(facade):
    """AllPaths of the facade (row f2) against the reference's candidates recorded in tests/golden/f2_vectors.json, through the
    same flat signature as the reference driver (names, exon chains, lengths, pre-mRNA flags, log numbers)."""
    import json
    from oracle import f2_py
    import test_f2_allpaths as t2
    ref_like = f2_py.ReferenceF2.__new__(f2_py.ReferenceF2)
    facade.facade_all_paths.restype = C.c_int

    class _L(object):
        ref_all_paths = facade.facade_all_paths
    ref_like.lib = _L
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "f2_vectors.json")))
    for ci, c in enumerate(golden):
        dot, nopre, maxc = t2._case(c)
        t2._same(ref_like.all_paths(dot, nopre, maxc), c["want"], ci)
    facade.facade_all_paths.argtypes = None
    assert facade.facade_all_paths(b"digraph g {", 0, 100, 0, 0, *([None] * 6), None, 0, None, 0, None) == -2  # throws, does not abort
