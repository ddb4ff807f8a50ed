#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_2gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tee gpurun_out/bench7_n1.log | tail -1 | cut -c1-400
for EX in p2p nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 5 --exchange $EX 2>&1 | tee gpurun_out/bench7_n2_$EX.log | tail -2 | cut -c1-600
done
