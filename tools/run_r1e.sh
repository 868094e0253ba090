set -x
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --pairs 200000 --steps 2 --warmup 1 > gpurun_out/bench2_r1e.json 2> gpurun_out/bench2_r1e.err; tail -5 gpurun_out/bench2_r1e.err; cat gpurun_out/bench2_r1e.json
