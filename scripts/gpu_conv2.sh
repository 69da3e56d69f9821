cd $GRAFT_REPO_ROOT
echo "=== default build: conv block tests"
timeout 600 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "conv_block" 2>&1 | tail -6
echo "=== default build: unet predict / prediction tests"
timeout 900 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "predict_matches or auto_runs or prediction_matches" 2>&1 | tail -8
echo "=== no-rebalance build: conv block + unet tests"
CT3D_LIB=$GRAFT_REPO_ROOT/3deecelltracker_b200/libct3d_norb.so timeout 900 python -m pytest tests/test_gpu_lcn_unet.py -m gpu -x -q -k "conv_block or predict_matches" 2>&1 | tail -6
echo "=== per-layer timing (default build)"
timeout 300 python scripts/conv_layers.py 38 tcgen05_classic tcgen05 2>&1 | tail -18
echo "=== watershed on named configs + pipeline tests"
timeout 900 python -m pytest tests/test_gpu_watershed.py tests/test_gpu_pipeline.py -x -q --durations=5 2>&1 | tail -15
timeout 300 python scripts/ws_time.py 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 3000 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
