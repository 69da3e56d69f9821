cd $GRAFT_REPO_ROOT
timeout 600 python scripts/tcz_debug.py 2>&1 | grep -v "tcgen05_split  " | awk '{ if ($0 ~ /max [0-9.]+e-0[78]/) n++; else print } END { print n " block cases below 1e-6" }' | tail -12
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py tests/test_gpu_spatial.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | grep -vE "^frame|^$" | tail -5 | cut -c1-250
bash scripts/gpu_tczlist.sh 2>&1 | tail -2 | cut -c1-1300
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-c3 2>gpurun_out/bench_q.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('ms_per_step','frames_per_s','serial_ms_per_step','stage_ms_per_step')}, d['e2e'], d['roofline']['frac'])"
grep -v "^frame" gpurun_out/bench_q.err | tail -3 | cut -c1-300
