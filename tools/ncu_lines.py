#!/usr/bin/env python
"""Top CUDA source lines by warp-stall samples from an .ncu-rep (needs -lineinfo and
--import-source on).  usage: tools/ncu_lines.py report.ncu-rep [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = {}
fname = func = None
hdr = None
seen_funcs = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1][:40]
        seen_funcs += 1
        continue
    if r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples")
        i_i = hdr.index("Instructions Executed")
        continue
    if hdr is None or seen_funcs > 1 and False:
        continue
    if r[0] not in ("", "-"):
        try:
            key = (fname, int(r[0]), r[1].strip()[:95])
            s, ins = float(r[i_s]), float(r[i_i])
        except Exception:
            continue
        a = agg.setdefault(key, [0.0, 0.0])
        a[0] += s
        a[1] += ins
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
print("total samples %.0f, total warp-instructions %.0f (all captured launches)" % (tot, toti))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / toti, k[0], k[1], k[2]))
