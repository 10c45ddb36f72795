"""A/B of the calibration pass (bayes_gibbs_set_calibration) on a synthetic configuration: per setting, upload time
(build + calibration + rebuild), step time of two runs, how well the corrected time model predicts unit durations and how
full the GPU is over the step.  Saves gpurun_out/unit_stats_c<iters>.npz for tools/schedule analysis.

    python tools/calib_ab.py [--config 2] [--iters 0,256,1024:BAYES_TAIL_PASSES=0,256:BAYES_TAIL_SLACK=1.1]

A variant is `iterations[:ENV=value]...`; the environment knobs are the experiment switches read by csrc/capi.cu.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from bayesembler_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--loci", type=int, default=0)
ap.add_argument("--iters", default="0,256,1024")
args = ap.parse_args()
batch = synth.SynthBatch(args.config, n_loci=args.loci or None)
d = batch.desc_array()
ref_occ = None
for variant in args.iters.split(","):
    parts = variant.split(":")
    it = int(parts[0])
    env = dict(p.split("=") for p in parts[1:])
    os.environ.update(env)
    ctx = bb.GibbsContext(0)
    ctx.set_calibration(it)
    ctx.submit(d)
    t0 = time.perf_counter()
    ctx.upload()
    ctx.sync()
    up_ms = 1e3 * (time.perf_counter() - t0)
    ctx.run()
    ctx.sync()
    t0 = time.perf_counter()
    ctx.run()
    ctx.run()
    ctx.sync()
    step_ms = 1e3 * (time.perf_counter() - t0) / 2
    shape, clock = ctx.unit_stats()
    t00 = clock[:, 0].min()
    start = (clock[:, 0] - t00).astype(np.float64) * 1e-6
    end = (clock[:, 1] - t00).astype(np.float64) * 1e-6
    dur = end - start
    pred = ctx.unit_predicted_s * 1e3
    ratio = dur / np.maximum(pred, 1e-9)
    warps = shape[:, 1] // 32
    ctx.download()
    ctx.sync()
    occ = ctx.fetch_all()["occ"]
    if ref_occ is None:
        ref_occ = occ
    same = bool(np.array_equal(ref_occ, occ))
    edges = np.linspace(0, end.max(), 21)
    fill = [(np.clip(np.minimum(end, b) - np.maximum(start, a), 0, None) * warps).sum() / (b - a) / 148 for a, b in zip(edges[:-1], edges[1:])]
    print("calibration %5d %s: upload %7.1f ms  step %7.1f ms  makespan %7.1f  longest unit %7.1f  warp-ms %.0f  launches %d  results identical %s"
          % (it, env, up_ms, step_ms, end.max(), dur.max(), (dur * warps).sum(), ctx.last_launches(), same))
    print("   measured/predicted duration: p5 %.2f p25 %.2f p50 %.2f p75 %.2f p95 %.2f   block sizes %s" % (
        *np.percentile(ratio, [5, 25, 50, 75, 95]), dict(zip(*[x.tolist() for x in np.unique(shape[:, 1], return_counts=True)]))))
    print("   resident warps/SM over the step: " + " ".join("%.0f" % o for o in fill))
    print("   host phases: %s" % ctx.phase_seconds())
    np.savez(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "unit_stats_c%s.npz" % variant.replace(":", "_").replace("=", "")), shape=shape,
             slab_h=(shape[:, 7] >> 8) & 255, smem_kb=shape[:, 7] >> 16, sm=clock[:, 2].astype(np.int32), start=start, end=end, pred=pred, feat=ctx.unit_features)
    ctx.close()
    for k in env:
        del os.environ[k]
