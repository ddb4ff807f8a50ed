#!/bin/bash
# round 2, GPU call 9 (4 GPUs): interleaved push, backward push vs pull
N=${1:-4}
mkdir -p gpurun_out
SWEEP_LEAN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    scripts/r2/mg_sweep.py --gpus $N --k 128 --steps 20 2>gpurun_out/c9_sweep_n${N}_k128.err | grep "^{" | tee gpurun_out/c9_sweep_n${N}_k128.jsonl | cut -c1-330
tail -3 gpurun_out/c9_sweep_n${N}_k128.err | cut -c1-300
