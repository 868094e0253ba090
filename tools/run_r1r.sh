set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
python bench.py > gpurun_out/bench_r1r_default.json 2> gpurun_out/bench_r1r_default.err; tail -3 gpurun_out/bench_r1r_default.err; cat gpurun_out/bench_r1r_default.json
T1K_NO_FAST=1 T1K_PAIR_OCC=3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1r_nofast_occ3.json 2> gpurun_out/bench_r1r_nofast.err; cat gpurun_out/bench_r1r_nofast_occ3.json
ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r1r -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r1r.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair -c 1 -o gpurun_out/prof_pair_r1r -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pair_r1r.log 2>&1
