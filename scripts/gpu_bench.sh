cd $GRAFT_REPO_ROOT
R=${1:-r1}
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; tail -c 2500 gpurun_out/bench_${R}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_${R}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${R}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prgls_kernel -s 15 -c 1 -o gpurun_out/prof_${R}_em python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_em_${R}.log 2>&1
