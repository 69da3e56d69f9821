cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_tcx -s 7 -c 7 -o gpurun_out/prof_tcx python scripts/tc_prof.py tcgen05 15 2 > gpurun_out/tcxprof.log 2>&1
tail -2 gpurun_out/tcxprof.log
ls -la gpurun_out/prof_tcx.ncu-rep
