# One round of GPU evidence (run through gpurun): parity tests, smoke, the default bench line + reference arm, the DRAM-traffic
# capture at the bench's launch size, the launch list, full ncu captures of k_assign / k_pair, host timing.  Outputs in gpurun_out/.
# Summaries for profiles/: tools/ncu_summary.py, ncu_src.py, ncu_stalls.py, ncu_traffic.py.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
( time timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_round_default.json 2> gpurun_out/bench_round_default.err; tail -3 gpurun_out/bench_round_default.err; cat gpurun_out/bench_round_default.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_round_reference.json 2> gpurun_out/bench_round_reference.err; cat gpurun_out/bench_round_reference.json
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_assign -c 1 --csv --log-file gpurun_out/traffic_round.csv python bench.py --pairs 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_traffic_round.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_round.csv python bench.py --pairs 200000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_round.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_round -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_round.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pair -c 1 -o gpurun_out/prof_pair_round -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pair_round.log 2>&1
T1K_TIMING=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_round_timing.json 2> gpurun_out/bench_round_timing.err; grep "t1k timing" gpurun_out/bench_round_timing.err | grep -v "assign: \(launch\|input\|store\|tail\)" | tail -22
