#!/usr/bin/env python
"""Side-by-side per-layer conv launch times (last U-Net batch) from ncu gpu__time_duration launch lists.
usage: launch_table.py a.csv b.csv ..."""
import csv
import sys

NAMES = ["d0a 1>8", "d0b 8>16", "d1a 16>16", "d1b 16>32", "d2a 32>32", "d2b 32>64", "u2a 64>64", "u2b 64>64",
         "u1a 128>32", "u1b 32>32", "u0a 64>16", "u0b 16>16", "o_m2 32>8", "o_m1 8>8"]


def convs(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    out, other = [], 0.0
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        if "conv3_" in r[ki]:
            out.append((r[ki].split("(")[0].replace("void ct::", "").replace("conv3_", ""), v))
        else:
            other += v
    return out[-14:] if len(out) % 14 == 0 else [("first_conv (separate kernel)", 0.0)] + out[-13:], other


cols = [convs(p) for p in sys.argv[1:]]
print("layer".ljust(12) + "".join(p.split("launches_")[-1][:-4].rjust(34) for p in sys.argv[1:]))
for i, n in enumerate(NAMES):
    print(n.ljust(12) + "".join(f"{c[0][i][0][:22]:>24}{c[0][i][1]:9.1f}u" for c in cols))
print("sum".ljust(12) + "".join(f"{sum(v for _, v in c[0]):33.1f}u" for c in cols))
