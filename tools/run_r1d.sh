set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --pairs 200000 --steps 2 --warmup 1 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -3 gpurun_out/bench_r1d.err; cat gpurun_out/bench_r1d.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --pairs 50000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r1d -f python bench.py --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair$ -c 1 -o gpurun_out/prof_pair_r1d -f python bench.py --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2_r1d.log 2>&1
nproc; free -g | head -2
