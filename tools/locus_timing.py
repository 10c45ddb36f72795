"""Per-locus latency / throughput probe (development tool, run on the GPU box).

    python tools/locus_timing.py [--config 2] [--chains 1,512,4096]

For a few representative loci of a synthetic configuration: time per Gibbs iteration when the locus runs alone
(latency of the dependent chain) and when n_chains copies fill the GPU (throughput).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from bayesembler_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--loci", type=int, default=2000)
    ap.add_argument("--chains", default="1,256,4096")
    ap.add_argument("--iters", type=int, default=4000)
    ap.add_argument("--T", default="2,3,5,8,12,20,30,200,600")
    args = ap.parse_args()
    batch = synth.SynthBatch(args.config, n_loci=args.loci)
    picks = []
    for T in [int(x) for x in args.T.split(",")]:
        idx = np.nonzero(batch.T == T)[0]
        if len(idx) == 0:
            idx = np.nonzero((batch.T >= T) & (batch.T <= 1.5 * T))[0]
        if len(idx):
            picks.append(int(idx[np.argsort(batch.R[idx])[len(idx) // 2]]))
    ctx = bb.GibbsContext(0)
    print("%6s %4s %5s %6s %7s %7s | %s" % ("locus", "T", "R", "nnz", "N_f", "chains", "us/iter(alone or per chain)  Miter/s  Mrow/s  Gdraw/s"))
    for i in picks:
        for nc in [int(c) for c in args.chains.split(",")]:
            d = batch.desc_array(n_chains=nc, subset=[i])
            d["burn_in"] = args.iters // 11
            d["samples"] = args.iters - args.iters // 11
            ctx.clear()
            ctx.submit(d)
            ctx.upload()
            ctx.run()
            ctx.sync()
            t0 = time.perf_counter()
            ctx.run()
            ctx.sync()
            dt = time.perf_counter() - t0
            it = args.iters
            nnz = int(batch.cell_off[i + 1] - batch.cell_off[i])
            print("%6d %4d %5d %6d %7d %7d | %10.2f %9.2f %8.1f %8.2f" % (
                i, batch.T[i], batch.R[i], nnz, batch.n_frag[i], nc, 1e6 * dt / it, it * nc / dt / 1e6,
                float(it) * nc * batch.R[i] / dt / 1e6, float(it) * nc * batch.n_frag[i] / dt / 1e9))
    ctx.close()


if __name__ == "__main__":
    main()
