cd $GRAFT_REPO_ROOT
timeout 300 python scripts/conv_layers.py 38 tcgen05_split planewalk_split 2>&1 | tail -16 | grep -E "d0b|d1a|u0b|o_m|sum"
