cd $GRAFT_REPO_ROOT
N=${1:-2}
R=${2:-r1c}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --workload c3 --shape 448 448 36 --verify --steps 2 --warmup 3 2>gpurun_out/c3v_n$N.err | tail -1 | cut -c1-900; grep -v "^\s*$\|OMP_NUM\|\*\*\*" gpurun_out/c3v_n$N.err | tail -5
timeout 600 $TR --master-port 29512 bench.py --gpus $N --workload c3 --steps 3 --warmup 3 2>gpurun_out/c3_n$N.err | tail -1 > gpurun_out/bench_${R}_c3_n$N.json; cat gpurun_out/bench_${R}_c3_n$N.json | cut -c1-1500; grep -v "^\s*$\|OMP_NUM\|\*\*\*" gpurun_out/c3_n$N.err | tail -3
timeout 300 $TR --master-port 29514 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-400
