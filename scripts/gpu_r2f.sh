cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "conv_block or predict_matches or prediction_matches" 2>&1 | tail -4
timeout 300 python scripts/conv_layers.py 38 tcgen05 2>&1 | tail -16
timeout 300 python scripts/tc_prof.py auto 38 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv3|first_conv|pool_|upsample|head_" -c 400 --csv --log-file gpurun_out/launches_unet_f.csv python scripts/tc_prof.py auto 38 2 > gpurun_out/ncu_unet.log 2>&1; tail -1 gpurun_out/ncu_unet.log
timeout 600 python scripts/tcx_timing.py 2>&1 | tail -12
