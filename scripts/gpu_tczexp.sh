cd $GRAFT_REPO_ROOT
timeout 400 python scripts/tcz_timing.py -DTZ_EXP_NOSTORE 2>&1 | tail -5
rm -rf 3deecelltracker_b200/csrc/build_timing
timeout 400 python scripts/tcz_timing.py -DTZ_EXP_NOEPI 2>&1 | tail -5
