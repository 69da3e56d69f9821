cd $GRAFT_REPO_ROOT
export CT3D_LIB=$GRAFT_REPO_ROOT/3deecelltracker_b200/libct3d_dev2.so
export CT3D_TCX_DW=16
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "split_fp16 or conv_block_tcgen05" 2>&1 | tail -8
timeout 300 python scripts/conv_layers.py 38 tcgen05 tcgen05_split 2>&1 | tail -16
export CT3D_TCX_DW=8
timeout 300 python scripts/conv_layers.py 38 tcgen05 tcgen05_split 2>&1 | tail -16
