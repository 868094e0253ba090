set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
T1K_TIMING=1 python bench.py --pairs 200000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -80 gpurun_out/bench_r1i.err; cat gpurun_out/bench_r1i.json
