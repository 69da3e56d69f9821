cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -22 | tee gpurun_out/pytest_r2c.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/ws_time.py 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ws_ -c 600 --csv --log-file gpurun_out/launches_ws.csv python scripts/ws_time.py > gpurun_out/ncu_ws.log 2>&1; tail -2 gpurun_out/ncu_ws.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 1800 gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
timeout 300 python scripts/em_time.py 2>&1 | tail -12
