"""Shape of the units a many-chains run of one locus produces (development aid)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb
from bayesembler_b200 import synth
locus, chains, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
batch = synth.SynthBatch(2, ids=[locus])
d = batch.desc_array(n_chains=chains)
d["burn_in"] = iters // 11
d["samples"] = iters - iters // 11
ctx = bb.GibbsContext(0)
ctx.submit(d); ctx.upload(); ctx.run(); ctx.sync()
t0 = time.perf_counter(); ctx.run(); ctx.sync(); dt = time.perf_counter() - t0
shape, clock = ctx.unit_stats()
dur = (clock[:, 1] - clock[:, 0]) * 1e-6
print("T=%d R=%d units=%d launches=%d shape0=%s h=%d wall %.1f ms, unit dur mean %.1f max %.1f ms, Miter/s %.1f" % (
    batch.T[0], batch.R[0], len(shape), ctx.last_launches(), shape[0, :7].tolist(), (shape[0, 7] >> 8) & 255, dt * 1e3, dur.mean(), dur.max(), iters * chains / dt / 1e6))
t0 = (clock[:, 0] - clock[:, 0].min()) * 1e-6
print("start times: max %.1f ms; resident est %.0f units" % (t0.max(), dur.sum() / (dt * 1e3)))
