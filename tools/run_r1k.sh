set -x
nvidia-smi -L | wc -l; nproc
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -5
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --pairs 500000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/scale_r1k_$n.json 2> gpurun_out/scale_r1k_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --pairs 500000 --steps 2 --warmup 1 > gpurun_out/scale_r1k_$n.json 2> gpurun_out/scale_r1k_$n.err
  fi
  tail -2 gpurun_out/scale_r1k_$n.err; cat gpurun_out/scale_r1k_$n.json
done
