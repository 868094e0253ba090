set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1y_default.json 2> gpurun_out/bench_r1y_default.err; tail -3 gpurun_out/bench_r1y_default.err; cat gpurun_out/bench_r1y_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1y.csv python bench.py --pairs 200000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r1y.log 2>&1
