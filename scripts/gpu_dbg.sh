cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py tests/test_gpu_spatial.py -m gpu -x -q -k "normalize or median or lcn or spatial or decomp" 2>&1 | grep -vE "^frame|^$" | tail -5 | cut -c1-250
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:box_|select_" -c 40 --csv --log-file gpurun_out/launches_lcn.csv python bench.py --steps 2 --no-cpu-baseline --no-c3 --no-compare > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_lcn.csv")) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h = rows[hdr]; agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    v = float(r[h.index("Metric Value")].replace(",", "")); u = r[h.index("Metric Unit")]
    agg[r[h.index("Kernel Name")][:40]].append(v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v))
for k, v in agg.items(): print(f"{k:40s} n={len(v)} avg {sum(v)/len(v):7.1f} us")
PY
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-c3 2>gpurun_out/bench_q.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('ms_per_step','frames_per_s','serial_ms_per_step','stage_ms_per_step')}, d['e2e']['frames_per_s'], d['roofline']['frac'], d['roofline_secondary']['lcn'])"
grep -v "^frame" gpurun_out/bench_q.err | tail -3 | cut -c1-300
