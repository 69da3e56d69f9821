# HBM-side evidence for the memory-bound kernels of one frame: bytes, time, DRAM / L2 throughput (few-pass ncu metrics).
cd $GRAFT_REPO_ROOT
R=${1:-r1e}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none -k "regex:pool_kernel|upsample_kernel|box_|select_hist|head_scatter|ffn_pair|sgemm_bn|knn_features|first_conv|predict_one_rep|greedy" -s 0 -c 2000 --csv --log-file gpurun_out/membound_${R}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-overlap > gpurun_out/ncu_membound_${R}.log 2>&1
tail -2 gpurun_out/ncu_membound_${R}.log; wc -l gpurun_out/membound_${R}.csv
