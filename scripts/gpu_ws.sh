cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_watershed.py tests/test_gpu_pipeline.py -x -q --durations=5 2>&1 | tail -25
timeout 300 python scripts/ws_time.py 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ws_ -c 400 --csv --log-file gpurun_out/launches_ws.csv python scripts/ws_time.py > gpurun_out/ncu_ws.log 2>&1; tail -2 gpurun_out/ncu_ws.log
timeout 600 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "conv_block or predict_matches or auto_runs" 2>&1 | tail -8
timeout 300 python scripts/conv_layers.py 38 tcgen05_classic tcgen05 2>&1 | tail -18
