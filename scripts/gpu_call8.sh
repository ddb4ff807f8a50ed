#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_2gpu.log | tail -25
python scripts/kbench.py --blocks 100 --iters 10 --ks 16,32,128 --variants 3 2>&1 | tee gpurun_out/kbench8.log | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tee gpurun_out/bench8_n1.log | tail -1 | cut -c1-300
for EX in p2p nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 5 --exchange $EX --no-cpu 2>&1 | tee gpurun_out/bench8_n2_$EX.log | tail -1 | cut -c1-300
done
