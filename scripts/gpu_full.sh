# Full GPU validation + bench + profiles for a round (run under gpurun, 1 GPU).
cd $GRAFT_REPO_ROOT
R=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; tail -c 3000 gpurun_out/bench_${R}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_${R}.err; tail -c 600 gpurun_out/bench_${R}_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${R}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv3_tc|first_conv" -s 84 -c 14 -o gpurun_out/prof_${R}_conv_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv_${R}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prgls_kernel -s 15 -c 1 -o gpurun_out/prof_${R}_em python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_em_${R}.log 2>&1
ls -la gpurun_out | tail -12
