set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --pairs 200000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -3 gpurun_out/bench_r1g.err; cat gpurun_out/bench_r1g.json
python bench.py --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1g_1m.json 2> gpurun_out/bench_r1g_1m.err; tail -3 gpurun_out/bench_r1g_1m.err; cat gpurun_out/bench_r1g_1m.json
ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r1g -f python bench.py --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1g.log 2>&1
