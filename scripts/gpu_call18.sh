#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 2>&1 | tee gpurun_out/bench18.log | tail -1 | cut -c1-250
python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tee gpurun_out/bench18_ref.log | tail -1 | cut -c1-250
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_final2.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_spmm_tiles -s 6 -c 2 -o gpurun_out/prof18_bench -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu18.log 2>&1
python scripts/kbench.py --blocks 100 --iters 10 --variants 3 2>&1 | tee gpurun_out/kbench18.log | tail -5
