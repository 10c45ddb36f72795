"""Where does a launch's time go?  Runs a synthetic configuration once (after a warm-up) and reports, from the
per-unit start/end stamps the kernels record (bayes_gibbs_unit_stats): makespan, sum of unit times, resident-unit
average, the longest units with their shapes and their per-iteration latency.

    python tools/schedule_report.py [--config 2] [--loci 20000] [--chains 1] [--top 12]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesembler_b200 as bb  # noqa: E402
from bayesembler_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--loci", type=int, default=0)
    ap.add_argument("--chains", type=int, default=1)
    ap.add_argument("--top", type=int, default=12)
    args = ap.parse_args()
    batch = synth.SynthBatch(args.config, n_loci=args.loci or None)
    d = batch.desc_array(n_chains=args.chains)
    ctx = bb.GibbsContext(0)
    ctx.submit(d)
    ctx.upload()
    ctx.run()
    ctx.sync()
    import time
    t0 = time.perf_counter()
    ctx.run()
    ctx.sync()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    ctx.run()
    ctx.run()
    ctx.sync()
    wall2_ms = 1e3 * (time.perf_counter() - t0) / 2
    print("wall clock of one run() + sync: %.1f ms; two runs back to back: %.1f ms each" % (wall_ms, wall2_ms))
    shape, clock = ctx.unit_stats()
    slab_h = (shape[:, 7] >> 8) & 255
    shape[:, 7] &= 1
    st = batch.stats(n_chains=args.chains)
    t0 = clock[:, 0].min()
    start = (clock[:, 0] - t0).astype(np.float64) * 1e-6
    end = (clock[:, 1] - t0).astype(np.float64) * 1e-6
    dur = end - start
    span = end.max()
    print("launches per run: %d" % ctx.last_launches())
    print("units %d  makespan %.1f ms  sum(unit time) %.1f ms  avg resident units %.0f (of %d SMs x 32)  draws/s %.3g" % (
        len(dur), span, dur.sum(), dur.sum() / span, len(np.unique(clock[:, 2])), st["draws"] / (span * 1e-3)))
    warps = shape[:, 1] // 32
    print("warp-time %.1f ms-warps -> avg resident warps/SM %.1f" % ((dur * warps).sum(), (dur * warps).sum() / span / 148))
    print("\nby class (wide, threads, mat_smem): units, first start, last end, mean/max duration ms, mean us/iter")
    keys = sorted({(int(s[0]), int(s[1]), int(s[7])) for s in shape})
    for k in keys:
        m = (shape[:, 0] == k[0]) & (shape[:, 1] == k[1]) & (shape[:, 7] == k[2])
        print("  %s  n=%5d  start %8.1f  end %8.1f  dur mean %8.1f max %8.1f  us/iter mean %7.1f" % (
            k, m.sum(), start[m].min(), end[m].max(), dur[m].mean(), dur[m].max(), (1e3 * dur[m] / shape[m, 5]).mean()))
    def show(idx, title):
        print("\n%s: unit# | dur ms | start ms | wide thr slots slabs h cells iters T0 | us/iter" % title)
        for i in idx:
            s = shape[i]
            print("  %5d | %8.1f | %8.1f | %d %4d %2d %4d %2d %6d %6d %4d | %7.1f" % (
                i, dur[i], start[i], s[0], s[1], s[2], s[3], slab_h[i], s[4], s[5], s[6], 1e3 * dur[i] / s[5]))
    show(np.argsort(-end)[:args.top], "last to finish")
    show(np.argsort(-dur)[:args.top], "longest")
    show(np.arange(min(args.top, len(dur))), "first in launch order")
    pred = ctx.unit_predicted_s * 1e3
    ratio = dur / np.maximum(pred, 1e-9)
    print("\ntime model: measured / predicted duration, by block size (median, 10-90 %%), overall corr %.3f" % np.corrcoef(pred, dur)[0, 1])
    for b in sorted(set(shape[:, 1].tolist())):
        m = shape[:, 1] == b
        print("  %4d threads  n=%5d  median %.2f  (%.2f - %.2f)" % (b, m.sum(), np.median(ratio[m]), np.percentile(ratio[m], 10), np.percentile(ratio[m], 90)))
    np.savez(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "unit_stats.npz"), shape=shape, slab_h=slab_h,
             start=start, end=end, pred=pred, feat=ctx.unit_features)
    # utilisation timeline: resident warps over time in 20 buckets
    edges = np.linspace(0, span, 21)
    occ = []
    for a, b in zip(edges[:-1], edges[1:]):
        ov = np.clip(np.minimum(end, b) - np.maximum(start, a), 0, None)
        occ.append((ov * warps).sum() / (b - a) / 148)
    print("\nresident warps/SM over time (20 buckets): " + " ".join("%.0f" % o for o in occ))
    ctx.close()


if __name__ == "__main__":
    main()
