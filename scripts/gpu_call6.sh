#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python scripts/kbench.py --blocks 100 --iters 10 --variants 3,19,35,67 2>&1 | tee gpurun_out/kbench6.log | tail -22
python scripts/kbench.py --blocks 1000 --iters 5 --ks 128,16 --variants 3 2>&1 | tee gpurun_out/kbench6_big.log | tail -3
python bench.py --steps 20 --warmup 5 2>&1 | tee gpurun_out/bench6.log | tail -2
