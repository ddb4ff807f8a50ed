#!/bin/bash
# single GPU: full parity suite, narrow-k tile A/B, final bench, ncu launch list + full captures for profiles/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_full.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for BT in 1 0; do python scripts/kbench.py --blocks 100 --iters 10 --ks 16,32 --variants 3 --big-tiles $BT 2>&1 | grep '"k"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('big_tiles=$BT k', d['k'], 'ms', d['ms_flushed'], 'frac', d['frac_of_peak'])"; done | tee gpurun_out/kbench14.log
python bench.py --steps 20 --warmup 5 2>&1 | tee gpurun_out/bench14.log | tail -1 | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench14_ref.log | tail -1 | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_final.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_spmm_tiles -s 6 -c 2 -o gpurun_out/prof14_bench -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu14.log 2>&1
ls -la gpurun_out | grep -E "prof14|launches_bench_final"
