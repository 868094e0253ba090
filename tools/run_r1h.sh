set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --pairs 200000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; tail -3 gpurun_out/bench_r1h.err; cat gpurun_out/bench_r1h.json
ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r1h -f python bench.py --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair -c 1 -o gpurun_out/prof_pair_r1h -f python bench.py --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2_r1h.log 2>&1
