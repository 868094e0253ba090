set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for occ in 3 2; do
T1K_ASSIGN_OCC=$occ python bench.py --pairs 200000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1p_occ$occ.json 2> gpurun_out/bench_r1p_occ$occ.err; tail -2 gpurun_out/bench_r1p_occ$occ.err
done
ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r1p -f python bench.py --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1p.log 2>&1
