set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
( time python bench.py > gpurun_out/bench_r1q_default.json 2> gpurun_out/bench_r1q_default.err ); tail -3 gpurun_out/bench_r1q_default.err; cat gpurun_out/bench_r1q_default.json
ncu --set full --clock-control none --import-source on -k regex:k_assign -c 1 -o gpurun_out/prof_assign_r1q -f python bench.py --pairs 50000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r1q.log 2>&1
