#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.log | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-e2e 2>&1 | tee gpurun_out/bench20.log | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3), 'GF', round(d['value']), 'roof', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline'])"
