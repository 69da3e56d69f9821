cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2i.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("ms_per_step", "frames_per_s", "frames_per_s_without_watershed", "stage_ms_per_step", "serial_ms_per_step")}, d["roofline"]["frac"], d["e2e"]["frames_per_s"])
PY
tail -3 gpurun_out/bench_r2i.err
