set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -3 gpurun_out/bench_r2h.err; cat gpurun_out/bench_r2h.json
