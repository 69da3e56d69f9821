cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py tests/test_gpu_spatial.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ('ms_per_step','serial_ms_per_step','stage_ms_per_step','frames_per_s')}, d['roofline']['achieved'], d['e2e']['frames_per_s'])
PY
tail -3 gpurun_out/bench_quick.err
