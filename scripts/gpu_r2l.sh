cd $GRAFT_REPO_ROOT
export CT3D_LIB=$GRAFT_REPO_ROOT/3deecelltracker_b200/libct3d_dev.so
timeout 900 python -m pytest tests/test_gpu_watershed.py tests/test_gpu_pipeline.py tests/test_gpu_lcn_unet.py -m gpu -x -q -k "not named_configs or watershed" 2>&1 | tail -4
timeout 300 python scripts/ws_time.py 2>&1 | tail -2
for lag in 1 2 3; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 --ws-lag $lag > gpurun_out/bench_r2l_$lag.json 2> gpurun_out/bench_r2l.err; python - $lag <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_r2l_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("lag", sys.argv[1], {k: d[k] for k in ("ms_per_step", "frames_per_s", "frames_per_s_without_watershed", "stage_ms_per_step", "serial_ms_per_step")}, d["roofline"]["frac"], d["e2e"]["frames_per_s"])
PY
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 --tiles-per-batch 75 > gpurun_out/bench_r2l_t75.json 2> gpurun_out/bench_r2l.err; python - <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_r2l_t75.json").read().strip().splitlines()[-1])
print("t75", {k: d[k] for k in ("ms_per_step", "frames_per_s", "frames_per_s_without_watershed", "stage_ms_per_step", "serial_ms_per_step")}, d["roofline"]["frac"], d["e2e"]["frames_per_s"])
PY
tail -3 gpurun_out/bench_r2l.err
