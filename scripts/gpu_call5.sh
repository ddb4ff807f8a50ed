#!/bin/bash
# ncu passes on the bench command (1 GPU): launch list + full captures of the two SpMM launches of a step
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_spmm_tiles -s 6 -c 2 -o gpurun_out/prof5_bench -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu5.log 2>&1
for K in 16 32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_tiles -s 3 -c 1 -o gpurun_out/prof5_k$K -f \
     python scripts/kbench.py --blocks 100 --iters 1 --ks $K --variants 3 > gpurun_out/ncu5_k$K.log 2>&1
done
python scripts/kbench.py --blocks 100 --iters 10 --variants 3 2>&1 | tee gpurun_out/kbench5.log | tail -6
ls -la gpurun_out | tail -12
