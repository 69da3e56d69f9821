cd $GRAFT_REPO_ROOT
timeout 600 python scripts/tcz_debug.py 2>&1 | grep -v "tcgen05_split  " | awk '{ if ($0 ~ /max [0-9.]+e-0[78]/) n++; else print } END { print n " block cases below 1e-6" }' | tail -8
timeout 300 python scripts/conv_layers.py 38 tcgen05_split planewalk_split 2>&1 | tail -16 | grep -E "d0b|d1a|u0b|o_m|sum"
CT3D_TCZ_ALL=1 bash scripts/gpu_tczlist.sh 2>&1 | tail -1 | cut -c1-1300
