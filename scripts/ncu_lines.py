#!/usr/bin/env python
"""Attribute the stall samples / executed instructions of an ncu report to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <source.cu> <mangled-kernel-substring> [top] [demangled-substring] [nth launch]

ncu's CSV source page is SASS-only, so the SASS offsets are joined with `nvdisasm -g` line annotations of a cubin
built from the same source with the library's flags.
"""
import csv
import io
import os
import re
import subprocess
import sys
import collections

rep, src, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
kern_ncu = sys.argv[5] if len(sys.argv) > 5 else kern          # demangled-name substring for the ncu side
nth = int(sys.argv[6]) if len(sys.argv) > 6 else -1            # which matching launch (-1: all)
cub = "/tmp/_ncu_lines.cubin"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                "--expt-relaxed-constexpr", "-cubin", "-o", cub, src], check=True, stderr=subprocess.DEVNULL)
sass = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
# offset -> line for the wanted function
line_of, cur, infn = {}, None, False
for ln in sass.splitlines():
    if ln.startswith("\t.section\t.text."):
        infn = kern in ln
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.defaultdict(lambda: [0, 0])
stalls = collections.defaultdict(lambda: collections.Counter())
h, base, use = None, None, False
for r in rows:
    if r and r[0] == "Kernel Name":
        use = kern_ncu in r[1]
        if use:
            seen = globals().get("seen", -1) + 1
            globals()["seen"] = seen
            use = nth < 0 or seen == nth
        base = None
        continue
    if r and r[0] == "Address":
        h = r
        continue
    if not use or h is None or len(r) != len(h):
        continue
    d = dict(zip(h, r))
    addr = int(d["Address"], 16)
    if base is None:
        base = addr
    key = line_of.get(addr - base)
    s, ie = int(d["# Samples"]), int(d["Instructions Executed"])
    agg[key][0] += s
    agg[key][1] += ie
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
            stalls[key][k[6:]] += int(v)
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
text = open(src).read().splitlines()
print(f"samples {tot}  warp-instructions {toti}")
for key, (s, ie) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    srcline = ""
    if key and key[0] == os.path.basename(src) and key[1] <= len(text):
        srcline = text[key[1] - 1].strip()[:90]
    st = ",".join(f"{k}:{v}" for k, v in stalls[key].most_common(3))
    print(f"{s / tot:6.3f} inst {ie / toti:6.3f}  {key}  [{st}]  {srcline}")
