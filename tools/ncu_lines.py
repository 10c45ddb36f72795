"""Summarise an ncu report per CUDA source line: instructions executed, samples, avg active threads.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""
hdr = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples")
        i_inst = hdr.index("Instructions Executed")
        i_thr = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    if r[2] != "-":  # SASS row below a source line
        continue
    try:
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()])
        a[0] += int(r[i_samp] or 0)
        a[1] += int(r[i_inst] or 0)
        a[2] += int(r[i_thr] or 0)
    except (ValueError, IndexError):
        pass
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
print("%-16s %5s %7s %7s %6s  %s" % ("file", "line", "samp%", "inst%", "lanes", "source"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %5d %6.2f%% %6.2f%% %6.1f  %s" % (key[0], key[1], 100.0 * a[0] / tot_s, 100.0 * a[1] / tot_i,
                                                 a[2] / a[1] if a[1] else 0, a[3][:110]))
