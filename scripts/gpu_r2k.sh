cd $GRAFT_REPO_ROOT
export CT3D_LIB=$GRAFT_REPO_ROOT/3deecelltracker_b200/libct3d_dev.so
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py tests/test_gpu_watershed.py tests/test_gpu_correction.py tests/test_gpu_pipeline.py -m gpu -x -q -k "not named_configs or watershed" 2>&1 | tail -4
timeout 300 python scripts/ws_time.py 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ws_ -c 200 --csv --log-file gpurun_out/launches_ws_k.csv python scripts/ws_time.py > gpurun_out/ncu_ws.log 2>&1; tail -1 gpurun_out/ncu_ws.log
timeout 300 python scripts/tc_prof.py auto 38 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv3|first_conv|pool_|upsample|head_" -c 400 --csv --log-file gpurun_out/launches_unet_k.csv python scripts/tc_prof.py auto 38 2 > gpurun_out/ncu_unet.log 2>&1; tail -1 gpurun_out/ncu_unet.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2k.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("ms_per_step", "frames_per_s", "frames_per_s_without_watershed", "stage_ms_per_step", "serial_ms_per_step")}, d["roofline"]["frac"], d["e2e"]["frames_per_s"])
PY
tail -3 gpurun_out/bench_r2k.err
