# last check of the tree: all GPU tests, smoke, the bench as the driver runs it (both arms)
cd $GRAFT_REPO_ROOT
R=${1:-r2f}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; tail -c 600 gpurun_out/bench_${R}.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_${R}.err; tail -c 300 gpurun_out/bench_${R}_reference.json
tail -3 gpurun_out/bench_${R}.err
