#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python scripts/kbench.py --blocks 100 --iters 10 2>&1 | tee gpurun_out/kbench3.log
for V in 0 3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm -s 3 -c 1 -o gpurun_out/prof3_v${V}_k128 -f \
     python scripts/kbench.py --blocks 100 --iters 1 --ks 128 --variants $V > gpurun_out/ncu3_v${V}.log 2>&1
done
