"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, draw for draw.

Both sides draw from the same site-keyed Philox streams (DESIGN.md "Replay"); integer results must be identical and
floating point must agree to 1e-9 relative (north_star asks for 1e-6).  Both sampling modes of the product are held
to this bar: "replay" (the reference's order of draws; oracle mode SITE, itself pinned against the reference's code
in SEQ mode) and "fast" (the production mode: splitting trees + pooled uniform words; oracle mode FAST).
"""
import numpy as np
import pytest

from oracle.oracle_py import FAST, SITE, random_locus

ORACLE_MODE = {"replay": SITE, "fast": FAST}
MODES = ["replay", "fast"]

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _theta_for(rng, loc):
    th = rng.dirichlet(np.ones(loc.T) * 0.7)
    if loc.T > 3:
        th[rng.choice(loc.T, size=loc.T // 3, replace=False)] = 0.0
    P = np.zeros((loc.R, loc.T))
    for i in range(loc.R):
        s, e = loc.row_ptr[i], loc.row_ptr[i + 1]
        P[i, loc.col_idx[s:e]] = loc.val[s:e]
    bad = (P * th).sum(1) <= 0
    for i in np.nonzero(bad)[0]:
        th[loc.col_idx[loc.row_ptr[i]]] = rng.random() + 0.01
    return th / th.sum()


@pytest.mark.parametrize("alpha", [1.0, 0.3, 1.7, 12.0, 5000.5])
def test_gamma_variates(lib, oracle, alpha):
    key = 0xA5A5A5A512345678
    g = lib.step_variates("gamma", alpha, 0.0, key, 20000)
    o = oracle.gamma(alpha, SITE, key, 20000)
    np.testing.assert_allclose(g, o, rtol=1e-12)


@pytest.mark.parametrize("n,p", [(1, 0.3), (19, 0.5), (25, 0.07), (40, 0.93), (300, 0.2), (5000, 0.5), (100000, 0.999), (7, 1e-9)])
def test_binomial_variates(lib, oracle, n, p):
    key = 0x0123456789ABCDEF
    g = lib.step_variates("binomial", n, p, key, 20000).astype(np.int64)
    o = oracle.binomial(n, p, SITE, key, 20000)
    assert np.array_equal(g, o)


@pytest.mark.parametrize("rng_n", [2, 3, 7, 1000, 4000000000])
def test_uniform_int_variates(lib, oracle, rng_n):
    key = 77
    g = lib.step_variates("uniform_int", rng_n, 0.0, key, 20000).astype(np.int64)
    o = oracle.uniform_int(rng_n, SITE, key, 20000).astype(np.int64) & 0xFFFFFFFF
    assert np.array_equal(g, o)


SHAPES = [(2, 5, 30), (3, 40, 60), (10, 150, 300), (31, 400, 100), (33, 300, 2000), (100, 700, 50), (257, 200, 500), (1500, 90, 40),
          (4096, 40, 25)]


@pytest.mark.parametrize("T,R,mc", SHAPES)
def test_step_init_assignment(lib, oracle, T, R, mc):
    rng = np.random.default_rng(T * 1000 + R)
    loc = random_locus(rng, T, R, max_count=mc)
    key = int(rng.integers(1, 2 ** 62))
    g = lib.step_init_assignment(loc, key)
    o, _ = oracle.init_assignment(loc, SITE, key)
    assert g.sum() == loc.n_frag
    assert np.array_equal(g, o)


@pytest.mark.parametrize("T,R,mc", SHAPES)
def test_step_generate_assignment(lib, oracle, T, R, mc):
    rng = np.random.default_rng(T * 77 + R)
    loc = random_locus(rng, T, R, max_count=mc)
    key = int(rng.integers(1, 2 ** 62))
    for it in (0, 5, 123456):
        th = _theta_for(rng, loc)
        g = lib.step_generate_assignment(loc, th, key, it)
        o, _ = oracle.generate_assignment(loc, th, SITE, key, it)
        assert g.sum() == loc.n_frag
        assert np.array_equal(g, o)


@pytest.mark.parametrize("T,R,mc", SHAPES)
def test_step_assignment_fast_mode(lib, oracle, T, R, mc):
    """initAssignment and generateAssignment of the production mode (wide layout, candidate masks, splitting trees in
    global memory, pool blocks computed by the rows): the same counts as the oracle's restatement of that mode."""
    rng = np.random.default_rng(T * 13 + R)
    loc = random_locus(rng, T, R, max_count=mc)
    key = int(rng.integers(1, 2 ** 62))
    g = lib.step_assignment(loc, None, key, mode="fast")
    o, _ = oracle.init_assignment(loc, FAST, key)
    assert g.sum() == loc.n_frag
    assert np.array_equal(g, o)
    for it in (0, 7, 654321):
        th = _theta_for(rng, loc)
        g = lib.step_assignment(loc, th, key, it, mode="fast")
        o, _ = oracle.generate_assignment(loc, th, FAST, key, it)
        assert g.sum() == loc.n_frag
        assert np.array_equal(g, o)
    # and the replay mode through the same entry point
    th = _theta_for(rng, loc)
    assert np.array_equal(lib.step_assignment(loc, th, key, 3, mode="replay"), oracle.generate_assignment(loc, th, SITE, key, 3)[0])


def test_step_assignment_reports_a_row_without_candidates(lib):
    """theta == 0 on every candidate of a row: the reference asserts (assignmentGenerator.cpp:284), the library reports."""
    rng = np.random.default_rng(4)
    loc = random_locus(rng, 6, 20, max_count=30, single_frac=0.0)
    th = np.zeros(6)
    th[0] = 1.0
    dense = loc.dense()
    if not ((dense[:, 0] == 0).any()):
        pytest.skip("every row has a cell in column 0")
    for mode in MODES:
        with pytest.raises(lib.BayesError, match="non-zero probability"):
            lib.step_assignment(loc, th, 5, 0, mode=mode)


@pytest.mark.parametrize("T,gamma", [(2, 1.0), (7, 1.0), (32, 0.5), (33, 1.0), (100, 2.5), (1000, 1.0)])
def test_step_generate_expression(lib, oracle, T, gamma):
    rng = np.random.default_rng(T)
    key = int(rng.integers(1, 2 ** 62))
    for it in (0, 9):
        counts = (rng.random(T) < 0.5) * rng.integers(1, 500, T)
        counts[rng.integers(T)] = 17
        n_plus = int((counts > 0).sum())
        for b in sorted({n_plus, min(T, n_plus + 1), min(T, n_plus + 5), T}):
            g = lib.step_generate_expression(counts, b, gamma, key, it)
            o, order, _ = oracle.generate_expression(counts, b, gamma, SITE, key, it)
            assert np.array_equal(g > 0, o > 0)
            assert (g > 0).sum() == b
            np.testing.assert_allclose(g, o, rtol=RTOL)
            assert abs(g.sum() - 1) < 1e-12


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("T,R,mc,gamma", [(2, 6, 40, 1.0), (3, 30, 60, 1.0), (8, 120, 400, 1.0), (12, 60, 100, 0.5),
                                          (30, 500, 80, 1.0), (40, 200, 3000, 2.0), (100, 300, 50, 1.0), (300, 700, 40, 1.0)])
def test_full_chain_replay(lib, oracle, T, R, mc, gamma, mode):
    rng = np.random.default_rng(T * 31 + R)
    loc = random_locus(rng, T, R, max_count=mc, gamma=gamma)
    seed = int(rng.integers(1, 2 ** 62))
    burn, samples = 60 + 3 * T, 600 + 30 * T
    total = burn + samples
    ctx = lib.GibbsContext(0, mode=mode)
    descs = lib.make_desc_array([loc], [burn], [samples], [seed])
    ctx.submit(descs)
    ctx.set_trace(0, 0, total)
    ctx.upload(); ctx.run(); ctx.download(); ctx.sync()
    g = ctx.fetch(0)
    tr = ctx.fetch_trace()
    info = ctx.locus_info(0)
    ctx.close()
    o = oracle.run_sampler(loc, burn, samples, ORACLE_MODE[mode], lib.chain_key(seed, 0), trace_iters=total)
    assert info["pi"] == o["pi"]
    assert np.array_equal(tr["init_counts"], o["init_counts"])
    assert np.array_equal(tr["b_trace"], o["b_trace"])
    assert np.array_equal(tr["counts_trace"], o["counts_trace"])
    np.testing.assert_allclose(tr["theta_trace"], o["theta_trace"], rtol=RTOL, atol=0)
    assert np.array_equal(g["occ"], o["occ"])
    np.testing.assert_allclose(g["mean"], o["mean"], rtol=RTOL)
    np.testing.assert_allclose(g["m2"], o["m2"], rtol=1e-7, atol=1e-9 * np.abs(o["mean"]).max() ** 2)
    assert g["final_counts"].sum() == loc.n_frag
    assert np.array_equal(g["final_counts"], o["counts_trace"][-1])


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("calibration", [0, 40, 100000])
def test_batch_mixed_loci_and_chains(lib, oracle, calibration, mode):
    """calibration: the pass of bayes_gibbs_upload that times the first iterations of every unit and re-forms the batch
    (off / shorter than every chain / longer than every chain) must not change a single draw."""
    rng = np.random.default_rng(2024)
    shapes = [(1, 4, 30), (5, 50, 100), (2, 3, 10), (20, 90, 40), (1, 1, 5), (9, 33, 25), (64, 150, 60), (3, 1000, 30)]
    loci = [random_locus(rng, T, R, max_count=mc) for (T, R, mc) in shapes]
    seeds = [int(s) for s in rng.integers(1, 2 ** 62, len(loci))]
    burn = [50 + 5 * l.T for l in loci]
    samples = [10 * b for b in burn]
    chains = [1, 3, 2, 1, 2, 4, 1, 2]
    ctx = lib.GibbsContext(0, mode=mode)
    assert ctx.get_mode() == mode
    ctx.set_calibration(calibration)
    descs = lib.make_desc_array(loci, burn, samples, seeds, n_chains=chains)
    ctx.run_batch(descs)
    for i, loc in enumerate(loci):
        for c in range(chains[i]):
            g = ctx.fetch(i, c)
            o = oracle.run_sampler(loc, burn[i], samples[i], ORACLE_MODE[mode], lib.chain_key(seeds[i], c))  # own read cover per chain
            assert o["pi"] == ctx.chain_info(i, c)["pi"]
            assert np.array_equal(g["occ"], o["occ"]), (i, c)
            np.testing.assert_allclose(g["mean"], o["mean"], rtol=RTOL)
            assert g["final_counts"].sum() == loc.n_frag
    ctx.close()


def test_error_paths_and_call_order(lib):
    """Status codes instead of the reference's asserts; calls out of order are refused."""
    import ctypes as C
    rng = np.random.default_rng(1)
    loc = random_locus(rng, 4, 10)
    L = lib.load_library()
    ctx = lib.GibbsContext(0)
    with pytest.raises(lib.BayesError, match="upload first"):
        ctx.run()
    bad = random_locus(rng, 4, 10)
    bad.col_idx[1] = 77
    with pytest.raises(lib.BayesError, match="ascending"):
        ctx.submit(lib.make_desc_array([bad], [5], [50], [1]))
    assert L.bayes_gibbs_num_loci(ctx.h) == 0                      # nothing of a failed submit is kept
    too_wide = random_locus(rng, 5000, 3, density=0.001)
    with pytest.raises(lib.BayesError, match="4096"):
        ctx.submit(lib.make_desc_array([too_wide], [5], [50], [1]))
    with pytest.raises(lib.BayesError, match="samples"):
        ctx.submit(lib.make_desc_array([loc], [5], [0], [1]))
    no_length = random_locus(rng, 4, 10)
    no_length.efflen[2] = 0.0
    with pytest.raises(lib.BayesError, match="effective length"):
        ctx.submit(lib.make_desc_array([no_length], [5], [50], [1]))
    ctx.submit(lib.make_desc_array([loc], [5], [50], [1]))
    ctx.upload()
    with pytest.raises(lib.BayesError, match="clear"):
        ctx.submit(lib.make_desc_array([loc], [5], [50], [1]))
    with pytest.raises(lib.BayesError, match="before"):
        ctx.set_trace(0, 0, 5)
    with pytest.raises(lib.BayesError, match="download"):
        ctx.fetch(0)
    ctx.run(); ctx.download(); ctx.sync()
    assert ctx.fetch(0)["final_counts"].sum() == loc.n_frag
    assert L.bayes_gibbs_fetch(ctx.h, 5, 0, None, None, None, None) == -1
    assert L.bayes_gibbs_fetch(ctx.h, 0, 3, None, None, None, None) == -1
    h = C.c_void_p()
    assert L.bayes_gibbs_create(99, C.byref(h)) == -1              # no such device
    ctx.close()


def test_empty_batch_and_single_candidate_only(lib):
    ctx = lib.GibbsContext(0)
    ctx.run_batch(lib.make_desc_array([], [], [], []))
    assert ctx.last_launches() == 0 and ctx.num_jobs() == 0
    from oracle.oracle_py import Locus
    one = Locus(1, [0, 1, 2], [0, 0], [1.0, 1.0], [7, 5], [1500.0], 1000000)
    ctx.clear()
    ctx.run_batch(lib.make_desc_array([one], [10], [100], [3]))
    g = ctx.fetch(0)
    assert ctx.last_launches() == 0 and g["occ"][0] == 100 and g["m2"][0] == 0
    assert g["mean"][0] == (12 * 1e9) / (1000000 * 1500.0)
    ctx.close()


def test_concurrent_submit(lib, oracle):
    """bayes_gibbs_submit from several host threads at once (the reference's workers each own one graph, assembler.cpp:1420-1445):
    every caller gets its own ids and every locus its own chain."""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(77)
    loci = [random_locus(rng, int(rng.integers(2, 14)), int(rng.integers(5, 60)), max_count=40) for _ in range(48)]
    seeds = [int(s) for s in rng.integers(1, 2 ** 40, len(loci))]
    burn, samples = 40, 400
    ctx = lib.GibbsContext(0)
    descs = [lib.make_desc_array(loci[i:i + 4], [burn] * 4, [samples] * 4, seeds[i:i + 4]) for i in range(0, len(loci), 4)]
    with ThreadPoolExecutor(6) as ex:
        firsts = list(ex.map(ctx.submit, descs))
    assert sorted(firsts) == list(range(0, len(loci), 4))
    ctx.upload()
    ctx.run()
    ctx.download()
    ctx.sync()
    for first, base in zip(firsts, range(0, len(loci), 4)):
        for j in range(4):
            g = ctx.fetch(first + j)
            o = oracle.run_sampler(loci[base + j], burn, samples, FAST, lib.chain_key(seeds[base + j], 0))  # default mode
            assert np.array_equal(g["occ"], o["occ"]), (first, j)
    ctx.close()


@pytest.mark.parametrize("n_cta", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("T,R,mc", [(70, 90, 300), (300, 700, 40), (1000, 400, 120), (2500, 150, 60)])
def test_cluster_rows_full_chain(lib, oracle, monkeypatch, T, R, mc, n_cta):
    """Loci with more than 32 candidates in the production mode (gibbs_rows.cu): rows dealt over the CTAs of a thread-block
    cluster, replicated E-step, partial counts summed through distributed shared memory.  Whatever the cluster size, every
    draw, every count and every theta of the whole chain equal the oracle's."""
    monkeypatch.setenv("BAYES_FORCE_CTAS", str(n_cta))
    rng = np.random.default_rng(T * 7 + R)
    loc = random_locus(rng, T, R, max_count=mc)
    seed = int(rng.integers(1, 2 ** 62))
    burn, samples = 40, 360
    total = burn + samples
    ctx = lib.GibbsContext(0, mode="fast")
    ctx.submit(lib.make_desc_array([loc], [burn], [samples], [seed]))
    ctx.set_trace(0, 0, total)
    ctx.upload(); ctx.run(); ctx.download(); ctx.sync()
    g = ctx.fetch(0)
    tr = ctx.fetch_trace()
    ctx.close()
    o = oracle.run_sampler(loc, burn, samples, FAST, lib.chain_key(seed, 0), trace_iters=total)
    assert np.array_equal(tr["init_counts"], o["init_counts"])
    assert np.array_equal(tr["b_trace"], o["b_trace"])
    assert np.array_equal(tr["counts_trace"], o["counts_trace"])
    np.testing.assert_allclose(tr["theta_trace"], o["theta_trace"], rtol=RTOL, atol=0)
    assert np.array_equal(g["occ"], o["occ"])
    np.testing.assert_allclose(g["mean"], o["mean"], rtol=RTOL)
    np.testing.assert_allclose(g["m2"], o["m2"], rtol=1e-7, atol=1e-9 * np.abs(o["mean"]).max() ** 2)
    assert np.array_equal(g["final_counts"], o["counts_trace"][-1])


@pytest.mark.parametrize("warps", [64, 128, 256])
def test_cluster_bundle_full_chain(lib, oracle, monkeypatch, warps):
    """A latency-critical bundle (loci with up to 32 candidates) spread over a thread-block cluster of 2, 4 or 8 CTAs of 32 warps
    (gibbs_fast.cu): tiles and splitting trees dealt over the CTAs, partial counts summed through distributed shared memory.
    Same chain as the oracle's, whatever the cluster size; a several-slot bundle rides along."""
    monkeypatch.setenv("BAYES_FORCE_WARPS", str(warps))
    rng = np.random.default_rng(warps)
    loci = [random_locus(rng, 30, 900, max_count=4000), random_locus(rng, 5, 60, max_count=300), random_locus(rng, 5, 40, max_count=50),
            random_locus(rng, 5, 80, max_count=900)]
    seeds = [int(s) for s in rng.integers(1, 2 ** 62, len(loci))]
    burn, samples = 50, 450
    total = burn + samples
    ctx = lib.GibbsContext(0, mode="fast")
    ctx.submit(lib.make_desc_array(loci, [burn] * len(loci), [samples] * len(loci), seeds))
    ctx.set_trace(0, 0, total)
    ctx.upload(); ctx.run(); ctx.download(); ctx.sync()
    tr = ctx.fetch_trace()
    shape, _ = ctx.unit_stats()
    assert np.all(shape[:, 1] == 1024)  # blocks of 32 warps, the cluster on top
    for i, loc in enumerate(loci):
        g = ctx.fetch(i)
        o = oracle.run_sampler(loc, burn, samples, FAST, lib.chain_key(seeds[i], 0), trace_iters=total if i == 0 else 0)
        if i == 0:
            assert np.array_equal(tr["init_counts"], o["init_counts"])
            assert np.array_equal(tr["b_trace"], o["b_trace"])
            assert np.array_equal(tr["counts_trace"], o["counts_trace"])
            np.testing.assert_allclose(tr["theta_trace"], o["theta_trace"], rtol=RTOL, atol=0)
        assert np.array_equal(g["occ"], o["occ"]), i
        np.testing.assert_allclose(g["mean"], o["mean"], rtol=RTOL)
        np.testing.assert_allclose(g["m2"], o["m2"], rtol=1e-7, atol=1e-9 * np.abs(o["mean"]).max() ** 2)
        assert g["final_counts"].sum() == loc.n_frag
    ctx.close()


@pytest.mark.parametrize("mode", MODES)
def test_full_run_at_the_candidate_limit(lib, oracle, mode):
    """T = 4096 (kMaxT) with a 40-90 KB matrix, the whole chain, both modes: the per-candidate state alone is ~190 KB of shared memory,
    so everything else has to stay in global memory / L2 -- decided on the unit's whole footprint, never a launch that does not fit."""
    rng = np.random.default_rng(4096)
    T = 4096
    loc = random_locus(rng, T, 48, max_count=90, density=0.045)
    assert 40_000 <= 10 * loc.nnz <= 110_000
    seed = int(rng.integers(1, 2 ** 62))
    burn, samples = 6, 40
    ctx = lib.GibbsContext(0, mode=mode)
    ctx.submit(lib.make_desc_array([loc], [burn], [samples], [seed]))
    ctx.set_trace(0, 0, burn + samples)
    ctx.upload(); ctx.run(); ctx.download(); ctx.sync()
    g, tr = ctx.fetch(0), ctx.fetch_trace()
    ctx.close()
    o = oracle.run_sampler(loc, burn, samples, ORACLE_MODE[mode], lib.chain_key(seed, 0), trace_iters=burn + samples)
    assert np.array_equal(tr["b_trace"], o["b_trace"])
    assert np.array_equal(tr["counts_trace"], o["counts_trace"])
    np.testing.assert_allclose(tr["theta_trace"], o["theta_trace"], rtol=RTOL, atol=0)
    assert np.array_equal(g["occ"], o["occ"])
    np.testing.assert_allclose(g["mean"], o["mean"], rtol=RTOL)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("T", [33, 200])
def test_wide_locus_whose_rows_all_fold(lib, oracle, mode, T):
    """A locus with more than 32 candidates in which every row has a single cell: all fragments are folded into base counts, the
    unit has no rows at all (no matrix, no pool, no trees) and the chain is the E-step alone."""
    rng = np.random.default_rng(T)
    loc = random_locus(rng, T, 40, max_count=50, single_frac=1.0)
    seed = int(rng.integers(1, 2 ** 62))
    ctx = lib.GibbsContext(0, mode=mode)
    ctx.run_batch(lib.make_desc_array([loc], [20], [200], [seed]))
    g = ctx.fetch(0)
    ctx.close()
    o = oracle.run_sampler(loc, 20, 200, ORACLE_MODE[mode], lib.chain_key(seed, 0))
    assert np.array_equal(g["occ"], o["occ"])
    np.testing.assert_allclose(g["mean"], o["mean"], rtol=RTOL)
    assert g["final_counts"].sum() == loc.n_frag
