set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
T1K_TIMING=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1v_timing.json 2> gpurun_out/bench_r1v_timing.err; grep "t1k timing" gpurun_out/bench_r1v_timing.err | tail -70; cat gpurun_out/bench_r1v_timing.json
