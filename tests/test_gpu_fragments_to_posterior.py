"""Fragments -> posterior through the product only (likelihood builder on the host, sampler on the GPU), checked against the
two oracles chained the same way (oracle/f1_oracle.c -> oracle/bayes_oracle.c): the widened path of SURVEY.md 8f row f1."""
import numpy as np
import pytest

from oracle import f1_py
from oracle.oracle_py import FAST, Locus

pytestmark = pytest.mark.gpu


def test_fragments_to_posterior(lib, oracle):
    import subprocess, os
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"), "oracle"])
    orc_f1 = f1_py.OracleF1()
    rng = np.random.default_rng(21)
    raws = [f1_py.random_raw_locus(rng, n_exons=int(rng.integers(3, 9)), n_cand=int(rng.integers(2, 9)), n_frag=int(rng.integers(200, 1500)),
                                   strand_plus=bool(i % 2)) for i in range(8)]
    n_total = float(sum(r.n_frag for r in raws))
    loci, ref_loci, burns = [], [], []
    for raw in raws:
        model = lib.SequencingModel(200.0, 20.0, int(raw.cand_length.max()))
        g = lib.build_collapsed_map(raw, model)
        o = orc_f1.collapsed_map(raw)
        assert g["return_code"] == o["return_code"] == 0
        loci.append(lib.Locus(g["T"], g["row_ptr"], g["col_idx"], g["val"], g["counts"], g["efflen"], n_total))
        # the oracle's dense matrix as CSR
        rp, ci, va = [0], [], []
        for i in range(o["R"]):
            nz = np.nonzero(o["P"][i])[0]
            ci += nz.tolist()
            va += o["P"][i, nz].tolist()
            rp.append(len(ci))
        ref_loci.append(Locus(o["T"], rp, ci, va, o["counts"], o["efflen"], n_total))
        burns.append(100 + 6 * int(g["complexity"].max()))  # assembler.cpp:1598 with a shortened base
    samples = [10 * b for b in burns]
    seeds = list(range(1000, 1000 + len(loci)))
    ctx = lib.GibbsContext(0)
    ctx.run_batch(lib.make_desc_array(loci, burns, samples, seeds))
    for i in range(len(loci)):
        got = ctx.fetch(i)
        want = oracle.run_sampler(ref_loci[i], burns[i], samples[i], FAST, lib.chain_key(seeds[i], 0))
        assert np.array_equal(got["occ"], want["occ"]), i
        np.testing.assert_allclose(got["mean"], want["mean"], rtol=1e-9)
    ctx.close()
