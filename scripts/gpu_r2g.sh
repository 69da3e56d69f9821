cd $GRAFT_REPO_ROOT
export CT3D_LIB=$GRAFT_REPO_ROOT/3deecelltracker_b200/libct3d_dev.so
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "split_fp16" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "not split_fp16 and not named_configs" 2>&1 | tail -8
timeout 300 python scripts/conv_layers.py 38 tcgen05 tcgen05_split 2>&1 | tail -16
timeout 300 python scripts/tc_prof.py auto 38 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv3|first_conv|pool_|upsample|head_" -c 400 --csv --log-file gpurun_out/launches_unet_g.csv python scripts/tc_prof.py auto 38 2 > gpurun_out/ncu_unet.log 2>&1; tail -1 gpurun_out/ncu_unet.log
