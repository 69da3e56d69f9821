# Final validation of the tree: all GPU tests, smoke, bench (both arms), launch list.
cd $GRAFT_REPO_ROOT
R=${1:-r1f}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; tail -c 2600 gpurun_out/bench_${R}.json | cut -c1-1800
timeout 600 python bench.py > gpurun_out/bench_${R}_default.json 2>> gpurun_out/bench_${R}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${R}_default.json')); print('default flags:', d['steps'], d['ms_per_step'], d['e2e']['frames_per_s'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_${R}.err; tail -c 300 gpurun_out/bench_${R}_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${R}.log 2>&1
tail -3 gpurun_out/bench_${R}.err
