# Whole-round validation + evidence (run under gpurun, 1 GPU): all GPU tests, smoke, bench (both arms), the ncu launch
# list of the bench command, ncu --set full of the 17 conv launches of one 38-tile batch and of one EM launch.
# usage: gpu_round.sh <tag>      then here: python scripts/summarize_profiles.py <tag>
cd $GRAFT_REPO_ROOT
R=${1:-r2}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; tail -c 3500 gpurun_out/bench_${R}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_${R}.err; tail -c 700 gpurun_out/bench_${R}_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/ncu_bench_${R}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv3_t|first_conv" -s 17 -c 17 -o gpurun_out/prof_${R}_conv_tc python scripts/tc_prof.py auto 38 2 > gpurun_out/ncu_conv_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prgls_kernel -s 3 -c 1 -o gpurun_out/prof_${R}_em python scripts/em_time.py > gpurun_out/ncu_em_${R}.log 2>&1
timeout 600 python scripts/em_time.py > gpurun_out/em_timings_${R}.jsonl 2>&1; tail -3 gpurun_out/em_timings_${R}.jsonl
timeout 300 python scripts/ws_time.py > gpurun_out/ws_time_${R}.txt 2>&1; tail -3 gpurun_out/ws_time_${R}.txt
tail -3 gpurun_out/bench_${R}.err
ls -la gpurun_out | tail -12
