set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json
ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r2a -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r2a.log 2>&1
