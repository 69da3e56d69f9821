cd $GRAFT_REPO_ROOT
export CT3D_LIB=$GRAFT_REPO_ROOT/3deecelltracker_b200/libct3d_dev.so
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "not named_configs" 2>&1 | tail -4
timeout 300 python scripts/conv_layers.py 38 tcgen05 tcgen05_split 2>&1 | tail -16
timeout 300 python scripts/tc_prof.py auto 38 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv3|first_conv|pool_|upsample|head_" -c 400 --csv --log-file gpurun_out/launches_unet_j.csv python scripts/tc_prof.py auto 38 2 > gpurun_out/ncu_unet.log 2>&1; tail -1 gpurun_out/ncu_unet.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_t -s 17 -c 17 -o gpurun_out/prof_r2j_conv python scripts/tc_prof.py auto 38 2 > gpurun_out/ncu_conv_r2j.log 2>&1; tail -2 gpurun_out/ncu_conv_r2j.log
ls -la gpurun_out/prof_r2j_conv.ncu-rep
