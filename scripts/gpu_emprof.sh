cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prgls_kernel -s 5 -c 1 -o gpurun_out/prof_em python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_em.log 2>&1
tail -2 gpurun_out/ncu_em.log
