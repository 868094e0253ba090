set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
T1K_TIMING=1 python bench.py --pairs 200000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; tail -22 gpurun_out/bench_r1j.err; cat gpurun_out/bench_r1j.json
python bench.py --pairs 1000000 --steps 2 --warmup 1 > gpurun_out/bench_r1j_1m.json 2> gpurun_out/bench_r1j_1m.err; tail -3 gpurun_out/bench_r1j_1m.err; cat gpurun_out/bench_r1j_1m.json
