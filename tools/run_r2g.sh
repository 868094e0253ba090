set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
( time timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r2g_default.json 2> gpurun_out/bench_r2g_default.err; tail -3 gpurun_out/bench_r2g_default.err; cat gpurun_out/bench_r2g_default.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2g_reference.json 2> gpurun_out/bench_r2g_reference.err; cat gpurun_out/bench_r2g_reference.json
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_assign -c 1 --csv --log-file gpurun_out/traffic_r2g.csv python bench.py --pairs 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_traffic_r2g.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --pairs 200000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r2g.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r2g -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r2g.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pair -c 1 -o gpurun_out/prof_pair_r2g -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pair_r2g.log 2>&1
T1K_TIMING=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2g_timing.json 2> gpurun_out/bench_r2g_timing.err; grep "t1k timing" gpurun_out/bench_r2g_timing.err | grep -v "assign: \(launch\|input\|store\|tail\)" | tail -22
