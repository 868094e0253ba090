set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -3 gpurun_out/bench_r2b.err; cat gpurun_out/bench_r2b.json
T1K_TUNE=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2b_nopf.json 2> gpurun_out/bench_r2b_nopf.err; cat gpurun_out/bench_r2b_nopf.json
ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r2b -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r2b.log 2>&1
