cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "conv_block or predict_matches or auto_runs" 2>&1 | tail -4
for E in tcgen05_classic tcgen05; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$E.csv python scripts/tc_prof.py $E 15 2 > gpurun_out/tcprof_$E.log 2>&1
done
python scripts/launch_table.py gpurun_out/launches_tcgen05_classic.csv gpurun_out/launches_tcgen05.csv 2>&1 | tail -17
