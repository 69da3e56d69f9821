cd $GRAFT_REPO_ROOT
for T in 38 75 25; do
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c3 --no-compare --tiles-per-batch $T 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tiles/batch', $T, d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['frames_per_s'], d['roofline']['frac'])"
done
