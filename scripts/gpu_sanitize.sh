# compute-sanitizer memcheck over the plane-walk kernel on small blocks (every instantiation: layers 2 / 11: 16->16, 12: 32->8, 13: 8->8, 1: 8->16)
cd $GRAFT_REPO_ROOT
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/tcz_debug.py 1 2 12 13 2>&1 | grep -vE "tcgen05_split  " | tail -25 | cut -c1-220
