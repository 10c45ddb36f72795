"""The other BASELINE.json configurations as parity cases (bench.py measures configs[1] only): loci of the synthetic
config 1 (the CPU-runnable case), config 3 (200-1000 candidates: wide units whose matrix is streamed from L2/HBM),
config 4 (heavy-tailed fragments per locus) and config 5 (many chains per locus) run through the C ABI with shortened
iteration counts and are compared draw for draw with the CPU oracle."""
import numpy as np
import pytest

from oracle.oracle_py import SITE

pytestmark = pytest.mark.gpu


def _check(lib, oracle, batch, idx, burn, samples, n_chains=1):
    descs = batch.desc_array(subset=idx, n_chains=n_chains)
    descs["burn_in"] = burn
    descs["samples"] = samples
    ctx = lib.GibbsContext(0)
    ctx.run_batch(descs)
    shape, _ = ctx.unit_stats()
    for j, i in enumerate(idx):
        loc = batch.locus(int(i))
        for c in range(n_chains):
            o = oracle.run_sampler(loc, burn, samples, SITE, lib.chain_key(int(batch.seeds[i]), c))
            g = ctx.fetch(j, c)
            assert np.array_equal(g["occ"], o["occ"]), (int(i), c)
            np.testing.assert_allclose(g["mean"], o["mean"], rtol=1e-9)
            assert g["final_counts"].sum() == loc.n_frag
    ctx.close()
    return shape


def test_config1_loci(lib, oracle):
    from bayesembler_b200 import synth
    batch = synth.SynthBatch(1)
    multi = np.nonzero(batch.T > 1)[0]
    idx = multi[np.argsort(-batch.R[multi])[[0, 3, 20, 60]]]  # includes the largest locus of the configuration
    _check(lib, oracle, batch, idx, 40, 400)


def test_config3_wide_streaming_loci(lib, oracle):
    from bayesembler_b200 import synth
    batch = synth.SynthBatch(3, n_loci=6)
    assert batch.T.min() >= 200 and batch.T.max() <= 1000
    idx = np.array([int(np.argmin(batch.T)), int(np.argmax(batch.T))])
    shape = _check(lib, oracle, batch, idx, 10, 60)
    assert np.all(shape[:, 0] == 1)              # wide units
    assert np.any((shape[:, 7] & 1) == 0)        # at least one streams its matrix instead of staging it in shared memory


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
