cd $GRAFT_REPO_ROOT
N=${1:-8}
R=${2:-r1d}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus $N --workload c3 --shape 448 448 36 --verify --steps 2 --warmup 3 2>gpurun_out/c3v_n$N.err | tail -1 | cut -c1-700; grep -v "^\s*$\|OMP_NUM\|\*\*\*" gpurun_out/c3v_n$N.err | tail -5
timeout 300 $TR --master-port 29512 bench.py --gpus $N --workload c3 --steps 5 --warmup 3 2>gpurun_out/c3_n$N.err | tail -1 > gpurun_out/bench_${R}_c3_n$N.json; cat gpurun_out/bench_${R}_c3_n$N.json | cut -c1-1400; grep -v "^\s*$\|OMP_NUM\|\*\*\*" gpurun_out/c3_n$N.err | tail -3
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/c1_n$N.err | tail -1 > gpurun_out/bench_${R}_n$N.json; cat gpurun_out/bench_${R}_n$N.json | cut -c1-1400; grep -v "^\s*$\|OMP_NUM\|\*\*\*" gpurun_out/c1_n$N.err | tail -3
