# GPU validation of the current tree: tests, smoke, bench (overlapped + serial), c3 on one GPU.
cd $GRAFT_REPO_ROOT
R=${1:-r1c}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; tail -c 2800 gpurun_out/bench_${R}.json; tail -5 gpurun_out/bench_${R}.err
timeout 600 python bench.py --workload c3 --shape 448 448 36 --verify --steps 2 --warmup 3 2>&1 | tail -3 | cut -c1-1200
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/bench_${R}_c3_n1.json 2>> gpurun_out/bench_${R}.err; tail -c 1500 gpurun_out/bench_${R}_c3_n1.json
