"""Full-size run of BASELINE.json configs[1] (20 000 loci, 20 M fragments, default sampler settings) through the C ABI:
size-independent properties of the domain for every locus, run-to-run determinism, and a draw-for-draw spot check of a
few loci against the CPU oracle (which needs seconds per locus, so it cannot check them all)."""
import numpy as np
import pytest

from oracle.oracle_py import FAST

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full_run(lib):
    from bayesembler_b200 import synth
    batch = synth.SynthBatch(2)
    descs = batch.desc_array()
    ctx = lib.GibbsContext(0)
    ctx.run_batch(descs)
    res = ctx.fetch_all()
    yield batch, descs, ctx, res
    ctx.close()


def test_full_size_invariants(full_run):
    batch, descs, ctx, res = full_run
    assert batch.n == 20000 and 19_000_000 < batch.n_frag.sum() < 21_000_000
    t_off = batch.t_off
    seg = np.repeat(np.arange(batch.n), batch.T)
    # every fragment is assigned to exactly one candidate after the last iteration (assignmentGenerator.cpp:363)
    assert np.array_equal(np.bincount(seg, weights=res["final_counts"], minlength=batch.n).astype(np.int64), batch.n_frag)
    samples = np.repeat(batch.samples, batch.T)
    occ = res["occ"]
    assert occ.min() >= 0 and np.all(occ <= samples)
    # |S+_expr| >= 1 in every sampling iteration: the inclusion counts of a locus add up to at least `samples`
    assert np.all(np.bincount(seg, weights=occ, minlength=batch.n) >= batch.samples)
    # T == 1: closed form (gibbsSampler.cpp:57-67)
    one = np.nonzero(batch.T == 1)[0]
    assert np.array_equal(occ[t_off[one]], batch.samples[one])
    want = batch.n_frag[one] * 1e9 / (np.trunc(batch.n_total) * batch.efflen[t_off[one]])
    np.testing.assert_allclose(res["mean"][t_off[one]], want, rtol=1e-15)
    assert np.all(res["m2"][t_off[one]] == 0)
    # Welford accumulators: finite, mean > 0 where included, M2 >= 0, nothing where never included
    assert np.all(np.isfinite(res["mean"])) and np.all(np.isfinite(res["m2"]))
    assert np.all(res["mean"][occ > 0] > 0) and np.all(res["m2"] >= 0)
    assert np.all(res["mean"][occ == 0] == 0) and np.all(res["m2"][occ <= 1] == 0)
    # a candidate that holds fragments at the end was in the last simplex; folded single-cell rows pin a floor
    assert ctx.last_launches() >= 1


def test_full_size_deterministic(full_run, lib):
    batch, descs, ctx, res = full_run
    ctx.run()          # same seeds -> same chains, whatever the block scheduler did
    ctx.download()
    ctx.sync()
    again = ctx.fetch_all()
    for k in res:
        assert np.array_equal(res[k], again[k]), k


def test_full_size_schedule_independent(full_run, lib):
    """The default run above went through the calibration pass (bayes_gibbs_set_calibration: automatic for a batch of this
    size); without it the units are sized and ordered differently and every result is the same."""
    batch, descs, ctx, res = full_run
    other = lib.GibbsContext(0)
    other.set_calibration(0)
    other.run_batch(descs)
    plain = other.fetch_all()
    other.close()
    for k in res:
        assert np.array_equal(res[k], plain[k]), k


def test_full_size_spot_check_against_oracle(full_run, lib, oracle):
    batch, descs, ctx, res = full_run
    rng = np.random.default_rng(11)
    small = np.nonzero((batch.T > 1) & (batch.T <= 12))[0]
    large = np.nonzero(batch.T > 12)[0]
    # the units that are special in the full batch: the longest chains (T >= 28: several warps, critical-path sizing,
    # calibrated re-shaping), by their cost cells x iterations
    cells = (batch.cell_off[1:] - batch.cell_off[:-1]).astype(np.float64)
    cost = cells * (batch.burn.astype(np.float64) + batch.samples)
    longest = [int(i) for i in np.argsort(-cost)[:3]]
    assert all(batch.T[i] >= 20 for i in longest)
    picks = longest + [int(i) for i in rng.choice(large, 3, replace=False)] + [int(i) for i in rng.choice(small, 4, replace=False)]
    for i in picks:
        o = oracle.run_sampler(batch.locus(i), int(batch.burn[i]), int(batch.samples[i]), FAST, lib.chain_key(int(batch.seeds[i]), 0))
        g = ctx.fetch(i)
        assert np.array_equal(g["occ"], o["occ"]), i
        np.testing.assert_allclose(g["mean"], o["mean"], rtol=1e-9)
        np.testing.assert_allclose(g["m2"], o["m2"], rtol=1e-6, atol=1e-9 * np.abs(o["mean"]).max() ** 2)
