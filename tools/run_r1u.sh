set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1u_occ3.json 2> gpurun_out/bench_r1u_occ3.err; tail -3 gpurun_out/bench_r1u_occ3.err; cat gpurun_out/bench_r1u_occ3.json
T1K_ASSIGN_OCC=4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1u_occ4.json 2> gpurun_out/bench_r1u_occ4.err; cat gpurun_out/bench_r1u_occ4.json
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_assign -c 1 --csv --log-file gpurun_out/traffic_r1u.csv python bench.py --pairs 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_traffic_r1u.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r1u -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r1u.log 2>&1
