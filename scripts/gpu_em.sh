cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_ffn_prgls.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['achieved'])"
