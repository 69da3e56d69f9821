# Parity of the conv engines + per-layer launch times of one 15-tile batch for classic vs stacked.
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "conv_block or predict_matches or auto_runs" 2>&1 | tail -12
for E in tcgen05_classic tcgen05 tcgen05_stacked; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$E.csv python scripts/tc_prof.py $E 15 2 > gpurun_out/tcprof_$E.log 2>&1
  tail -2 gpurun_out/tcprof_$E.log
done
python scripts/launch_table.py gpurun_out/launches_tcgen05_classic.csv gpurun_out/launches_tcgen05.csv gpurun_out/launches_tcgen05_stacked.csv 2>&1 | tail -40
