"""Turn the scratch outputs of tools/gpu_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.

    python tools/summarise_profile.py r1a

Writes  profiles/<tag>_bench.json      the bench lines as printed (b200 arm, reference arm, 2000-locus slice)
        profiles/<tag>_launches.csv    per-kernel launch count / total / mean device time of the ncu launch list
        profiles/<tag>_ncu_full.csv    selected counters of the `--set full` capture, one row per captured launch
        profiles/traffic.json          dram bytes per launch of the dominant kernel (read by bench.py -> roofline.traffic)
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


def launches(tag):
    p = os.path.join(OUT, tag + "_launches.csv")
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = {}
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        name = r["Kernel Name"]
        key = (name, r.get("Block Size", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values()) or 1.0
    out = os.path.join(PROF, tag + "_launches.csv")
    with open(out, "w") as f:
        f.write("kernel,block,launches,total_ms,mean_ms,share\n")
        for (name, blk), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s","%s",%d,%.3f,%.3f,%.4f\n' % (name, blk, a[0], a[1] / 1e6, a[1] / 1e6 / a[0], a[1] / total))
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
    # dram traffic per launch of the longest captured launch = the dominant kernel
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v = float(r[ix[name]].replace(",", ""))
        u = units[ix[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9,
                    "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1)
    top = max(body, key=lambda r: val(r, "gpu__time_duration.sum"))
    traffic = val(top, "dram__bytes_read.sum") + val(top, "dram__bytes_write.sum")
    json.dump({"tag": tag, "kernel": top[ix["Kernel Name"]], "grid": top[ix["Grid Size"]], "block": top[ix["Block Size"]],
               "dram_bytes_per_launch": traffic, "duration_ns": val(top, "gpu__time_duration.sum"),
               "command": "python bench.py --loci 2000 --steps 1 --warmup 1 --no-cpu-baseline (ncu --set full, 2000-locus slice of config 2)"},
              open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
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
    print("launch list:", launches(tag))
    print("ncu full   :", ncu_full(tag))


if __name__ == "__main__":
    main()
