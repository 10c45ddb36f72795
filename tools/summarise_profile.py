"""Turn the scratch outputs of tools/gpu_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.

    python tools/summarise_profile.py r1a

Writes  profiles/<tag>_bench.json      the bench lines as printed (b200 arm, reference arm, 2000-locus slice)
        profiles/<tag>_launches.csv    per-kernel launch count / total / mean device time of the ncu launch list
        profiles/<tag>_ncu_full.csv    selected counters of the `--set full` capture, one row per captured launch
        profiles/<tag>_bulk_ncu_full.csv  the same for the saturated single-launch capture (tools/ncu_bulk.sh)
        profiles/traffic.json          DRAM bytes of one bench step's sampler launches (read by bench.py -> roofline.traffic)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_blocks",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(tag, launches_per_step):
    """Per (kernel, block size): launch count, device time, DRAM bytes; plus the DRAM traffic of ONE bench step
    (= launches_per_step consecutive launches of the sampler kernels, the unit bench.py's roofline.traffic refers to)."""
    p = os.path.join(OUT, tag + "_launches.csv")
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    scale = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
             "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_launch = {}
    for r in rows:
        v = float(r["Metric Value"].replace(",", "")) * scale.get(r.get("Metric Unit", ""), 1)
        d = per_launch.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "block": r.get("Block Size", ""), "ns": 0.0, "dram": 0.0})
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["ns"] = v
        elif r["Metric Name"].startswith("dram__bytes"):
            d["dram"] += v
    agg = {}
    for d in per_launch.values():
        a = agg.setdefault((d["name"], d["block"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["ns"]
        a[2] += d["dram"]
    total = sum(a[1] for a in agg.values()) or 1.0
    out = os.path.join(PROF, tag + "_launches.csv")
    with open(out, "w") as f:
        f.write("kernel,block,launches,total_ms,mean_ms,share_of_device_time,dram_MB_per_launch\n")
        for (name, blk), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s","%s",%d,%.3f,%.3f,%.4f,%.3f\n' % (name, blk, a[0], a[1] / 1e6, a[1] / 1e6 / a[0], a[1] / total, a[2] / 1e6 / a[0]))
    ours = [d for _, d in sorted(per_launch.items()) if "gibbs_" in d["name"] or "stagger" in d["name"]]
    if launches_per_step and len(ours) >= launches_per_step:
        step = ours[:launches_per_step]
        json.dump({"tag": tag, "what": "DRAM read+write bytes summed over the %d sampler launches of one bench step (ncu launch list, "
                   "dram__bytes_read.sum + dram__bytes_write.sum per launch)" % launches_per_step,
                   "dram_bytes_per_launch": sum(d["dram"] for d in step),
                   "serialised_device_ms": sum(d["ns"] for d in step) / 1e6,
                   "command": "python bench.py --no-cpu-baseline (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum)"},
                  open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    return out


def ncu_full(tag):
    rep = os.path.join(OUT, tag + "_prof.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or h in ("Kernel Name", "Block Size", "Grid Size")]
    out = os.path.join(PROF, tag + "_ncu_full.csv")
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in body:
            w.writerow([r[i] for i in cols])
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(PROF, exist_ok=True)
    lines = {}
    for name in ("bench", "bench_ref", "small"):
        p = os.path.join(OUT, "%s_%s.json" % (tag, name))
        if os.path.exists(p):
            for l in open(p):
                l = l.strip()
                if l.startswith("{"):
                    lines[name] = json.loads(l)
    if lines:
        json.dump(lines, open(os.path.join(PROF, tag + "_bench.json"), "w"), indent=1)
    print("bench lines:", list(lines))
    per_step = 0
    if "bench" in lines and lines["bench"].get("steps"):
        per_step = int(lines["bench"]["gpu_launches"]) // int(lines["bench"]["steps"])
    print("launch list:", launches(tag, per_step))
    print("ncu full   :", ncu_full(tag))
    bulk = os.path.join(OUT, tag + "_bulk_prof.ncu-rep")
    if os.path.exists(bulk):
        global_tag = tag + "_bulk"
        print("ncu bulk   :", ncu_full(global_tag))


if __name__ == "__main__":
    main()
