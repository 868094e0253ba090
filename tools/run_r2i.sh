set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_r2i_default.json 2> gpurun_out/bench_r2i_default.err; tail -3 gpurun_out/bench_r2i_default.err; cat gpurun_out/bench_r2i_default.json
