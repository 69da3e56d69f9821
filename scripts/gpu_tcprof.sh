cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python scripts/tc_prof.py tcgen05 15 2 > gpurun_out/tcprof.log 2>&1
tail -2 gpurun_out/tcprof.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_tc -s 14 -c 14 -o gpurun_out/prof_conv_tc python scripts/tc_prof.py tcgen05 15 2 > gpurun_out/tcprof2.log 2>&1
tail -2 gpurun_out/tcprof2.log
