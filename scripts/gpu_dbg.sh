cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "predict_matches or prediction_matches or normalize or auto_runs" 2>&1 | grep -vE "^frame|^$" | tail -4 | cut -c1-250
bash scripts/gpu_tczlist.sh 2>&1 | tail -2 | head -1 | cut -c1-1300
