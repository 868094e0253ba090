set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --pairs 400000 > gpurun_out/bench_r1w_2gpu.json 2> gpurun_out/bench_r1w_2gpu.err; tail -3 gpurun_out/bench_r1w_2gpu.err; cat gpurun_out/bench_r1w_2gpu.json
python bench.py --gpus 1 --steps 2 --warmup 1 --pairs 400000 --no-cpu-baseline > gpurun_out/bench_r1w_1gpu.json 2> gpurun_out/bench_r1w_1gpu.err; cat gpurun_out/bench_r1w_1gpu.json
