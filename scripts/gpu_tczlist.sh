cd $GRAFT_REPO_ROOT
for ALL in 0 1; do
CT3D_TCZ_ALL=$ALL timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv3_t|first_conv" -s 17 -c 17 --csv --log-file gpurun_out/launches_tcz_$ALL.csv python scripts/tc_prof.py auto 38 2 > /dev/null 2>&1
python - $ALL <<'PY'
import csv, sys
rows = [r for r in csv.reader(open("gpurun_out/launches_tcz_%s.csv" % sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h = rows[hdr]
names = ["d0a", "d0b", "d1a", "d1b", "d2a", "d2b", "u2a", "u2b", "u1a up", "u1a skip", "u1b", "u0a up", "u0a skip", "u0b", "o_m2 up", "o_m2 skip", "o_m1"]
tot = 0; out = []
for n, r in zip(names, rows[hdr + 1:]):
    v = float(r[h.index("Metric Value")].replace(",", "")); u = r[h.index("Metric Unit")]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    tot += v
    out.append(f"{n} {r[h.index('Kernel Name')][11:20]} {v:.0f}")
print("TCZ_ALL=" + sys.argv[1], " | ".join(out), "| total", round(tot))
PY
done
