set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err
cat gpurun_out/bench_r1_b.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_direct.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_direct -s 40 -c 3 -o gpurun_out/prof_conv_direct python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prgls_kernel -s 5 -c 1 -o gpurun_out/prof_em python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_em.log 2>&1
ls -la gpurun_out
