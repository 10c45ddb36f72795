"""The BASELINE.json configurations as parity cases (bench.py measures their throughput): loci of the synthetic
config 1 (the CPU-runnable case), config 3 (200-1000 candidates: wide units whose matrix is streamed from L2/HBM),
config 4 (heavy-tailed fragments per locus) and config 5 (many chains per locus) run through the C ABI with shortened
iteration counts and are compared draw for draw with the CPU oracle."""
import numpy as np
import pytest

from oracle.oracle_py import FAST, SITE

ORACLE_MODE = {"replay": SITE, "fast": FAST}

pytestmark = pytest.mark.gpu


def _check(lib, oracle, batch, idx, burn, samples, n_chains=1, mode="fast"):
    descs = batch.desc_array(subset=idx, n_chains=n_chains)
    descs["burn_in"] = burn
    descs["samples"] = samples
    ctx = lib.GibbsContext(0, mode=mode)
    ctx.run_batch(descs)
    shape, _ = ctx.unit_stats()
    for j, i in enumerate(idx):
        loc = batch.locus(int(i))
        for c in range(n_chains):
            o = oracle.run_sampler(loc, burn, samples, ORACLE_MODE[mode], lib.chain_key(int(batch.seeds[i]), c))
            g = ctx.fetch(j, c)
            assert np.array_equal(g["occ"], o["occ"]), (int(i), c)
            np.testing.assert_allclose(g["mean"], o["mean"], rtol=1e-9)
            assert g["final_counts"].sum() == loc.n_frag
    ctx.close()
    return shape


@pytest.mark.parametrize("mode", ["fast", "replay"])
def test_config1_loci(lib, oracle, mode):
    from bayesembler_b200 import synth
    batch = synth.SynthBatch(1)
    multi = np.nonzero(batch.T > 1)[0]
    idx = multi[np.argsort(-batch.R[multi])[[0, 3, 20, 60]]]  # includes the largest locus of the configuration
    _check(lib, oracle, batch, idx, 40, 400, mode=mode)


@pytest.mark.parametrize("mode,ctas", [("fast", 0), ("fast", 1), ("replay", 0)])
def test_config3_wide_streaming_loci(lib, oracle, monkeypatch, mode, ctas):
    """ctas = 0: the library's own choice -- in the fast mode two loci alone on the GPU get a cluster of 16 CTAs each, every CTA with
    its share of the matrix resident in shared memory; ctas = 1: one CTA per locus (what a batch of hundreds of such loci gets), the
    largest matrices streamed from L2."""
    from bayesembler_b200 import synth
    if ctas:
        monkeypatch.setenv("BAYES_FORCE_CTAS", str(ctas))
    batch = synth.SynthBatch(3, n_loci=6)
    assert batch.T.min() >= 200 and batch.T.max() <= 1000
    idx = np.array([int(np.argmin(batch.T)), int(np.argmax(batch.T))])
    shape = _check(lib, oracle, batch, idx, 10, 60, mode=mode)
    assert np.all(shape[:, 0] == 1)              # wide units
    if mode == "fast" and ctas == 0:
        assert np.all((shape[:, 7] & 1) == 1)    # spread over a cluster: resident
    else:
        assert np.any((shape[:, 7] & 1) == 0)    # at least one streams its matrix instead of staging it in shared memory


@pytest.mark.parametrize("mode,pick", [("fast", [0, 9, 19, 29, 39]), ("replay", [0, 39])])
def test_config3_long_traces(lib, oracle, mode, pick):
    """Real config-3 loci, from the narrowest to the widest (T ~ 1000), for 2 200 iterations each with the full theta /
    counts / simplex-size traces: long enough for the simplex to thin out, so the per-row caches of the cells in play and
    the splitting trees run in the regime the full-length chains spend their time in."""
    from bayesembler_b200 import synth
    batch = synth.SynthBatch(3, n_loci=40)
    order = np.argsort(batch.T)
    burn, samples = 200, 2000
    for i in order[pick]:
        i = int(i)
        loc = batch.locus(i)
        descs = batch.desc_array(subset=[i])
        descs["burn_in"] = burn
        descs["samples"] = samples
        ctx = lib.GibbsContext(0, mode=mode)
        ctx.submit(descs)
        ctx.set_trace(0, 0, burn + samples)
        ctx.upload(); ctx.run(); ctx.download(); ctx.sync()
        g, tr = ctx.fetch(0), ctx.fetch_trace()
        ctx.close()
        o = oracle.run_sampler(loc, burn, samples, ORACLE_MODE[mode], lib.chain_key(int(batch.seeds[i]), 0), trace_iters=burn + samples)
        assert np.array_equal(tr["init_counts"], o["init_counts"]), i
        assert np.array_equal(tr["b_trace"], o["b_trace"]), i
        assert np.array_equal(tr["counts_trace"], o["counts_trace"]), i
        np.testing.assert_allclose(tr["theta_trace"], o["theta_trace"], rtol=1e-9, atol=0)
        assert np.array_equal(g["occ"], o["occ"]), i
        np.testing.assert_allclose(g["mean"], o["mean"], rtol=1e-9)


def test_config4_heavy_tail_loci(lib, oracle):
    from bayesembler_b200 import synth
    batch = synth.SynthBatch(4, n_loci=400)
    multi = np.nonzero(batch.T > 1)[0]
    idx = multi[np.argsort(-batch.n_frag[multi])[[0, 1, 50]]]
    _check(lib, oracle, batch, idx, 30, 300)


def test_config5_many_chains(lib, oracle):
    from bayesembler_b200 import synth
    batch = synth.SynthBatch(5, n_loci=40)
    multi = np.nonzero((batch.T > 1) & (batch.T <= 10))[0][:3]
    _check(lib, oracle, batch, multi, 30, 300, n_chains=8)
