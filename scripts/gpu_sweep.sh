cd $GRAFT_REPO_ROOT
for R in 1 2 4 8 16; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --reserve-sms $R > gpurun_out/bench_rs$R.json 2> gpurun_out/bench_rs$R.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_rs$R.json'))
print($R, {k:d[k] for k in ('ms_per_step','serial_ms_per_step','stage_ms_per_step')}, d['e2e']['frames_per_s'])
PY
done
