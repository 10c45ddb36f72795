"""Aggregate an ncu report's source page by code region (functions of rng.cuh / gibbs_kernel.cu): stall samples, warp
instructions and average active lanes per region.

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep
"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg, text = "", None, {}, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) <= max(i_s, i_i, i_t) or r[2] != "-":
        continue
    try:
        k = (cur, int(r[0]))
        a = agg.setdefault(k, [0, 0, 0])
        a[0] += int(r[i_s] or 0)
        a[1] += int(r[i_i] or 0)
        a[2] += int(r[i_t] or 0)
        text[k] = r[1].strip()
    except ValueError:
        pass

# region boundaries: lines that open a function in our sources
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bounds = {}
for f in ("rng.cuh", "gibbs_kernel.cu", "gibbs_fast.cu", "gibbs_rows.cu", "gibbs_common.cuh"):
    src = open(os.path.join(ROOT, "bayesembler_b200", "csrc", f)).read().splitlines()
    marks = []
    for i, line in enumerate(src, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:BAYES_HD(?:_NOINLINE)?|__device__|__global__|static|inline).*?\b([A-Za-z_0-9]+)\s*\(", line)
        if m and not line.startswith(" "):
            marks.append((i, m.group(1)))
        ms = re.match(r"^struct\s+([A-Za-z_0-9]+)\s*\{", line)  # member functions count for their struct
        if ms:
            marks.append((i, ms.group(1)))
    bounds[f] = marks


def region(f, l):
    if f not in bounds:
        return f
    name = f
    for (i, n) in bounds[f]:
        if i <= l:
            name = n
        else:
            break
    return name


reg = {}
for (f, l), a in agg.items():
    r = reg.setdefault(region(f, l), [0, 0, 0])
    for j in range(3):
        r[j] += a[j]
ts = sum(a[0] for a in reg.values()) or 1
ti = sum(a[1] for a in reg.values()) or 1
print("total stall samples %d, warp instructions %d, avg lanes %.1f" % (ts, ti, sum(a[2] for a in reg.values()) / ti))
print("%-28s %7s %7s %6s" % ("region", "samp%", "inst%", "lanes"))
for n, a in sorted(reg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %6.1f%% %6.1f%% %6.1f" % (n, 100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2] / a[1] if a[1] else 0))
