cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_ffn_prgls.py -m gpu -x -q 2>&1 | grep -vE "^frame|^$" | tail -4 | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c3 2>gpurun_out/bench_q.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('ms_per_step','frames_per_s','serial_ms_per_step','stage_ms_per_step')}, d['e2e']['frames_per_s'], d['roofline']['frac'])"
grep -v "^frame" gpurun_out/bench_q.err | tail -3 | cut -c1-300
