set -x
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --pairs 500000 --steps 2 --warmup 1 > gpurun_out/scale_r1l_2.json 2> gpurun_out/scale_r1l_2.err
tail -2 gpurun_out/scale_r1l_2.err; cat gpurun_out/scale_r1l_2.json
