cd $GRAFT_REPO_ROOT
for V in a b c d e f; do
CT3D_E2E_TRACE=1 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-c3 --no-compare 2>gpurun_out/tr_$V.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run', d['ms_per_step'], d['e2e']['frames_per_s'])"
grep "device-arm" gpurun_out/tr_$V.err | cut -c44-200
done
