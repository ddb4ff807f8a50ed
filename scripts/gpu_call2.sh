#!/bin/bash
# GPU call: parity tests, kernel sweep, ncu captures of the default SpMM kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_call2.csv &
SMI=$!
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python scripts/kbench.py --blocks 100 --iters 10 2>&1 | tee gpurun_out/kbench2.log
kill $SMI
for K in 16 128; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_tiles -s 3 -c 1 -o gpurun_out/prof_tiles_k$K -f \
     python scripts/kbench.py --blocks 100 --iters 1 --ks $K --variants 3 > gpurun_out/ncu_k$K.log 2>&1
done
ls -la gpurun_out
