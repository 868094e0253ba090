set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -3 gpurun_out/bench_r2c.err; cat gpurun_out/bench_r2c.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r2c -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r2c.log 2>&1
