# GPU test suite + smoke + short bench (run under gpurun, 1 GPU).  Usage: gpu_tests.sh <tag> [pytest -k expression]
cd $GRAFT_REPO_ROOT
R=${1:-r2}
K=${2:-}
if [ -n "$K" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$K" --durations=8 2>&1 | tail -25 | tee gpurun_out/pytest_${R}.log
else
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee gpurun_out/pytest_${R}.log
fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; tail -c 2500 gpurun_out/bench_${R}.json; tail -5 gpurun_out/bench_${R}.err
