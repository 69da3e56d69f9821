cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "predict_matches or prediction_matches or auto_runs" 2>&1 | tail -3
for T in 15 25 38 75; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tiles-per-batch $T > gpurun_out/bench_tpb$T.json 2> gpurun_out/bench_tpb$T.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_tpb$T.json'))
print($T, {k:d[k] for k in ('ms_per_step','serial_ms_per_step','stage_ms_per_step')}, d['e2e']['frames_per_s'])
PY
tail -2 gpurun_out/bench_tpb$T.err
done
