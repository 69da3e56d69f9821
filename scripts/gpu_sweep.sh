cd $GRAFT_REPO_ROOT
for R in 4 6 8 12; do
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c3 --no-compare --reserve-sms $R 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reserve', $R, d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['frames_per_s'], d['roofline']['frac'])"
done
