set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for occ in 4 5 6; do
T1K_ASSIGN_OCC=$occ python bench.py --pairs 200000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r1m_occ$occ.json 2> gpurun_out/bench_r1m_occ$occ.err; tail -2 gpurun_out/bench_r1m_occ$occ.err
done
