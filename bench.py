#!/usr/bin/env python
"""bench.py -- fragment-assignment draws/s of the per-locus Gibbs sampler (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C] [--loci n]

One "step" = one pass of the hot path over one batch of synthetic loci: every locus of the workload runs ALL of its
Gibbs iterations (burn-in 1000 + 60 T, samples 10 x burn-in; assembler.cpp:1598-1599).

Workload: ONE fixed workload at every N -- BASELINE.json configs[3] ("100M fragments over ~40k loci, cost-balanced locus
sharding at 1/2/4/8 B200"), the configuration the metric's "at 1/2/4/8 B200" is quoted on; it fits one GPU.  With
--gpus N the same 40 000 loci are sharded over the N ranks by the cost-balanced greedy partition (bayes_partition_lpt),
one process per GPU, no collective on the data path: STRONG scaling ("scaling": "strong"; the driver's v_N / (N v_1) is
then the strong-scaling efficiency of a fixed job).  `--config 2` etc. select another configuration; at N = 1 the line
also carries `other_configs`: configs 1, 2, 5 in full and a stated sample of config 3, at full iteration counts.

  value    draws/s with the packed loci already resident in HBM (kernel launches only, CUDA events on the launch stream)
  e2e      the same metric through the C ABI with HOST buffers: submit (pack, read cover, CDF tables) + H2D + kernels +
           D2H + fetch of every step inside the timed region; two contexts, double-buffered (the host prepares step k + 1
           while the kernels of step k run)
  roofline algorithmic bytes (12 nnz + 8 R + 24 T per iteration, + 40 T when sampling) / device time of the
           step's concurrent sampler launches, against the measured HBM copy bandwidth
  cpu_baseline / --impl reference: the UNMODIFIED reference sampler (oracle/_ref, built from /root/reference against
           the stand-in headers) or, where that .so is absent, the C restatement -- on the host cores of this box, over
           a cost-stratified sample of the same workload taken largest first like the reference's own queue
           (assembler.cpp:1886-1916).  Its inputs come from the repo's synthetic generator (libbayessynth.so: inputs
           only, no sampler code); its variate transforms are the stand-in's (oracle/orc_rng.h; Boost is absent).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fragment_assignment_draws_per_sec"
UNIT = "draws/s"
DEFAULT_CONFIG = 4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=DEFAULT_CONFIG)
    ap.add_argument("--loci", type=int, default=0, help="override the total number of loci (0 = the config's own)")
    ap.add_argument("--chains", type=int, default=0, help="chains per locus (0 = 64 for config 5, else 1)")
    ap.add_argument("--mode", default=None, choices=["fast", "replay"], help="sampling mode (default: the library's, fast)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end measurement (0 = min(steps, 5))")
    return ap.parse_args()


def workload_string(cfg, n_loci, n_chains, fragments):
    return ("config %d (BASELINE.json configs[%d]): %d loci x %d chain(s), %d fragments, default sampler settings "
            "(gamma=1, burn 1000+60T, samples 10x)" % (cfg, cfg - 1, n_loci, n_chains, fragments))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def locus_costs(batch):
    return (batch.burn.astype(np.float64) + batch.samples) * (batch.cell_off[1:] - batch.cell_off[:-1] + 4 * batch.R + 8 * batch.T)


def pick_sample(costs, k):
    """k loci spread evenly over the cost-sorted population (stratified by cost), most expensive first: the reference's
    workers take the graphs largest first (assembler.cpp:1886-1916), so no thread is left alone with a long locus at the end."""
    order = np.argsort(costs)
    k = min(k, len(order))
    pos = ((np.arange(k) + 0.5) * len(order) / k).astype(np.int64)
    return order[pos][::-1].copy()


def cpu_reference_rate(batch, sample_idx, n_threads):
    """Time the reference's own sampler on the host cores over a bounded sample; returns (draws/s, kind, seconds)."""
    from oracle import oracle_py
    loci = [batch.locus(int(i)) for i in sample_idx]
    seeds = (batch.seeds[sample_idx] & np.uint64(0x7FFFFFFF)).astype(np.uint32)
    burn, samples = batch.burn[sample_idx], batch.samples[sample_idx]
    draws = float(((burn.astype(np.int64) + samples) * batch.n_frag[sample_idx])[batch.T[sample_idx] > 1].sum())
    if os.path.exists(oracle_py.REF_SO):
        ref = oracle_py.Reference()
        secs, _ = ref.run_batch(loci, seeds, burn, samples, n_threads)
        return draws / secs, "reference", secs
    # the oracle port, one locus per worker thread (ctypes releases the GIL)
    orc = oracle_py.Oracle()
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:
        list(ex.map(lambda j: orc.run_sampler(loci[j], int(burn[j]), int(samples[j]), oracle_py.SEQ, int(seeds[j])), range(len(loci))))
    secs = time.perf_counter() - t0
    return draws / secs, "port", secs


def sized_sample(batch, costs, n_threads, target_seconds):
    """Pilot on a few loci, then size the stratified sample for about target_seconds of wall time."""
    multi = np.nonzero(batch.T > 1)[0]
    pilot = multi[pick_sample(costs[multi], max(4, n_threads))]
    rate, kind, secs = cpu_reference_rate(batch, pilot, n_threads)
    per_locus = secs / max(1, len(pilot))
    k = int(max(len(pilot), min(len(multi), target_seconds / max(per_locus, 1e-6))))
    return multi[pick_sample(costs[multi], k)]


def sample_note(sample, batch, cfg, n_threads, secs=None):
    s = "%d of %d loci of config %d, stratified by cost, taken largest first, full iteration counts, %d threads" % (
        len(sample), batch.n, cfg, n_threads)
    return s if secs is None else s + ", %.1f s wall" % secs


def run_reference_arm(args, cfg, total_loci, n_chains):
    from bayesembler_b200 import synth
    from oracle import oracle_py
    oracle_py.build(ref=True)
    n_threads = os.cpu_count() or 1
    batch = synth.SynthBatch(cfg, n_loci=total_loci, n_threads=n_threads)
    costs = locus_costs(batch)
    sample = sized_sample(batch, costs, n_threads, args.cpu_seconds)
    rates, secs_all = [], []
    kind = "port"
    for i in range(args.warmup + args.steps):
        rate, kind, secs = cpu_reference_rate(batch, sample, n_threads)
        if i >= args.warmup:
            rates.append(rate)
            secs_all.append(secs)
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs_all)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(cfg, batch.n, n_chains, int(batch.n_frag.sum()))},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": kind,
                             "sample": sample_note(sample, batch, cfg, n_threads) + " per step (a rate: draws of the sample / its wall time)",
                             "notes": "inputs from the repo's synthetic generator (libbayessynth.so, no sampler code); reference TUs compiled "
                                      "unmodified (-O3, asserts on) against stand-in Eigen/Boost headers whose variate transforms are "
                                      "oracle/orc_rng.h (Boost absent): inversion binomial below n*q = 32, BTRD above (switching at 11, about where Boost "
                                      "does, is 2 % slower on the CPU: profiles/cpu_binomial_switch.txt)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def time_resident(ctx, descs, steps, warmup, stream, flush, barrier, torch):
    """Device-resident throughput of the submitted batch: kernels only, CUDA events on the launch stream."""
    ctx.submit(descs)
    ctx.upload()
    ctx.sync()
    for _ in range(warmup):
        ctx.run()
    ctx.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed steps (costs ~0.05 ms of a step that takes hundreds)
            ctx.run()
        ev1.record(stream)
    ctx.sync()
    barrier()
    return ev0.elapsed_time(ev1), ctx.last_launches() * steps


def config3_sample(batch, k):
    """Config 3 at full iteration counts is minutes per locus on any hardware (150 000 - 670 000 iterations of a locus with
    hundreds of candidates): the bench line carries a stated sample -- k loci spread evenly over the candidate-count order."""
    order = np.argsort(batch.T)
    pos = ((np.arange(k) + 0.5) * len(order) / k).astype(np.int64)
    return order[pos]


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = args.config
    n_chains = args.chains or (64 if cfg == 5 else 1)

    from bayesembler_b200 import synth
    total_loci = args.loci or synth.CONFIGS[cfg][0]
    # host threads of this rank: pack / read cover / CDF tables of the end-to-end path.  Never fewer than 8: the ranks
    # do not prepare at the same instant for long, and a starved submit shows up in e2e directly.
    host_threads = max(min(8, os.cpu_count() or 1), (os.cpu_count() or 1) // max(1, world))

    # ------------------------------------------------------------------ reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, cfg, total_loci, n_chains)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    import bayesembler_b200 as bb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sampler has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # workload: total_loci loci whatever the number of ranks.  Every rank generates a contiguous block and costs it, the
    # costs are exchanged, the LPT partition is computed identically everywhere and each rank regenerates exactly the
    # loci assigned to it (the generator is counter-based: any locus is regenerable from its id).
    from bayesembler_b200 import shard
    lo, hi = rank * total_loci // world, (rank + 1) * total_loci // world
    block = synth.SynthBatch(cfg, n_loci=hi - lo, first=lo, n_threads=host_threads)
    descs = block.desc_array(n_chains=n_chains)
    cost_local = shard.desc_costs(descs)
    if world > 1:
        # blocks may differ by one locus: pad to a common length for the all-gather
        per = (total_loci + world - 1) // world
        padded = np.zeros(per)
        padded[:len(cost_local)] = cost_local
        gathered = shard.gather_costs(padded, dist, device="cuda").reshape(world, per)
        cost_all = np.concatenate([gathered[r, :(r + 1) * total_loci // world - r * total_loci // world] for r in range(world)])
        mine, load = shard.shard_loci(cost_all, world, rank)
        batch = synth.SynthBatch(cfg, ids=mine, n_threads=host_threads)
        descs = batch.desc_array(n_chains=n_chains)
        cost_local = cost_all[mine]
        shard_cost = [float(x) for x in load]
    else:
        batch = block
        shard_cost = [float(cost_local.sum())]
    st = batch.stats(n_chains=n_chains)

    stream = torch.cuda.Stream()
    ctx = bb.GibbsContext(local_rank, stream=stream.cuda_stream, host_threads=host_threads, mode=args.mode)
    mode = ctx.get_mode()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    # ---- device-resident throughput: kernels only
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms, launches = time_resident(ctx, descs, args.steps, args.warmup, stream, flush, barrier, torch)
    clocks = sampler.stop()
    ctx.download()
    ctx.sync()
    res = ctx.fetch_all()
    checksum = int(res["occ"].astype(np.int64).sum())

    # ---- end to end through the C ABI from host buffers
    h2d = int(descs.nbytes + batch.row_ptr.nbytes + batch.col.nbytes + batch.val.nbytes + batch.count.nbytes + batch.efflen.nbytes)
    d2h = int(res["occ"].nbytes + res["mean"].nbytes + res["m2"].nbytes + res["final_counts"].nbytes)
    # Two contexts, double-buffered, as a caller with a stream of batches would use the library: while the kernels of step k
    # run, the host prepares step k + 1 (fold, read covers, CDF tables, packing) and enqueues its H2D copies, launches and D2H
    # copies; then it waits for step k and reads its results.  Every step pays its own submit + H2D + kernels + D2H + fetch.
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 5))
    ctx2 = bb.GibbsContext(local_rank, host_threads=host_threads, mode=args.mode)
    pipe = [ctx, ctx2]

    def enqueue(c):
        c.clear()
        c.submit(descs)
        c.upload()
        c.run()
        c.download()

    def collect(c):
        c.sync()
        return c.fetch_all()

    for i in range(min(args.warmup, 1) * 2):
        enqueue(pipe[i % 2])
        collect(pipe[i % 2])
    barrier()
    t0 = time.perf_counter()
    outs = []
    for k in range(e2e_steps):
        enqueue(pipe[k % 2])
        if k > 0:
            outs.append(collect(pipe[(k - 1) % 2]))
    outs.append(collect(pipe[(e2e_steps - 1) % 2]))
    t1 = time.perf_counter()
    host_phases = ctx.phase_seconds()
    barrier()
    e2e_ms = 1e3 * (t1 - t0)
    for out in outs:
        assert int(out["occ"].astype(np.int64).sum()) == checksum  # same seeds -> same chains as the resident run
    ctx2.close()

    # ---- aggregate: max time over ranks, sum of work
    agg = torch.tensor([kernel_ms, e2e_ms, st["draws"], st["algo_bytes"], st["row_visits"], st["cell_visits"], st["loci"],
                        st["iterations"], float(launches), float(st["fragments"]), float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        kernel_ms, e2e_ms = float(mx[0]), float(mx[1])
        draws, algo_bytes, row_visits, cell_visits, loci, iterations = [float(x) for x in sm[2:8]]
        launches, fragments, h2d, d2h = int(sm[8]), int(sm[9]), int(sm[10]), int(sm[11])
    else:
        draws, algo_bytes, row_visits, cell_visits, loci, iterations = st["draws"], st["algo_bytes"], st["row_visits"], st["cell_visits"], st["loci"], st["iterations"]
        fragments = int(st["fragments"])

    line = None
    if rank == 0:
        secs = kernel_ms / 1e3
        value = draws * args.steps / secs
        peak, peak_src = measured_peaks()
        achieved = algo_bytes * args.steps / secs / 1e9 / world  # per GPU
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        traffic = None
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(cfg, int(loci), n_chains, fragments),
                       "sampling_mode": mode,
                       "sharding": "the one workload sharded over %d rank(s) by cost (LPT), no data-path collective" % world,
                       "shard_cost_imbalance": (max(shard_cost) / (sum(shard_cost) / len(shard_cost))) if shard_cost else 1.0,
                       "l2": "flushed between timed steps (256 MB write on the launch stream); each step re-stages every locus "
                             "from HBM into shared memory (%.0f MB of inputs over all ranks)" % (h2d / 1e6)},
            "loci_per_sec": loci * args.steps / secs, "row_visits_per_sec": row_visits * args.steps / secs,
            "cell_visits_per_sec": cell_visits * args.steps / secs, "iterations_per_sec": iterations * args.steps / secs,
            "e2e": {"value": draws * e2e_steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps, "loci_per_sec": loci * e2e_steps / (e2e_ms / 1e3),
                    "host_phases_s_rank0": host_phases, "host_threads_per_rank": host_threads,
                    "pipelining": "two contexts, double-buffered: step k+1 is prepared, uploaded and enqueued on the host while the kernels "
                                  "of step k run; every step pays its own submit + H2D + kernels + D2H + fetch"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "note": "dominant kernel = gibbs_fast_bundle_kernel (sampling mode fast; gibbs_bundle_kernel in replay mode; loci with more "
                                 "than 32 candidates -- config 3 -- run in gibbs_rows_kernel); one step "
                                 "issues its unit classes as concurrent launches of that kernel, so 'per launch' here means per step: achieved = "
                                 "algorithmic bytes of the step (sum over jobs of iters*(12 nnz + 8 R + 24 T) + samples*40 T, f64/u32 "
                                 "reference-equivalent CSR) / CUDA-event time of the step, per GPU; traffic = DRAM bytes summed over the same "
                                 "launches (profiles/traffic.json). The loci are shared-memory resident, so DRAM traffic is about one read per "
                                 "locus and the kernel is issue/latency-bound, not HBM-bound (DESIGN.md section 4)"},
            "clocks": clocks,
        }

    # ---- the other configurations, full iteration counts (N = 1 only: they are parity / coverage cases, not the scaled job)
    if rank == 0 and world == 1 and not args.no_other_configs and not args.loci:
        others = {}
        peak, _ = measured_peaks()
        for oc in (1, 2, 5, 3):
            if oc == cfg:
                continue
            try:
                chains = 64 if oc == 5 else 1
                ob = synth.SynthBatch(oc, n_threads=host_threads)
                subset, note = None, "all %d loci" % ob.n
                if oc == 3:
                    subset = config3_sample(ob, 16)
                    note = "%d of %d loci, evenly spread over the candidate-count order (T = %s), every chain at its full length" % (
                        len(subset), ob.n, ",".join(str(int(t)) for t in ob.T[subset]))
                od = ob.desc_array(n_chains=chains, subset=subset)
                ost = ob.stats(n_chains=chains, subset=subset)
                ctx.clear()
                w, k = (1, 2) if oc != 3 else (0, 1)
                ms, _ = time_resident(ctx, od, k, w, stream, flush, barrier, torch)
                s = ms / 1e3 / k
                others["config%d" % oc] = {
                    "workload": workload_string(oc, int(ost["loci"]), chains, int(ost["fragments"])), "loci_measured": note,
                    "ms_per_step": ms / k, "steps": k, "warmup": w, "value": ost["draws"] / s, "unit": UNIT,
                    "loci_per_sec": ost["loci"] / s, "cell_visits_per_sec": ost["cell_visits"] / s,
                    "roofline_frac": ost["algo_bytes"] / s / 1e9 / peak}
                if oc == 2 and not args.no_cpu_baseline:
                    # configs[1] was the N = 1 headline of round 1: its own CPU reference rate, for comparison across rounds
                    try:
                        from oracle import oracle_py
                        oracle_py.build(ref=True)
                        n_threads = os.cpu_count() or 1
                        smp = sized_sample(ob, locus_costs(ob), n_threads, 8.0)
                        rate, kind, secs_cpu = cpu_reference_rate(ob, smp, n_threads)
                        others["config2"]["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": n_threads, "kind": kind,
                                                             "sample": sample_note(smp, ob, 2, n_threads, secs_cpu)}
                    except Exception as e:
                        others["config2"]["cpu_baseline"] = {"failed": repr(e)}
            except Exception as e:  # reported, never allowed to take the headline down
                others["config%d" % oc] = {"failed": repr(e)}
        line["other_configs"] = others

    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                from oracle import oracle_py
                oracle_py.build(ref=True)
                n_threads = os.cpu_count() or 1
                sample = sized_sample(batch, cost_local, n_threads, args.cpu_seconds)
                rate, kind, secs_cpu = cpu_reference_rate(batch, sample, n_threads)
                line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": n_threads, "kind": kind,
                                        "sample": sample_note(sample, batch, cfg, n_threads, secs_cpu)}
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
