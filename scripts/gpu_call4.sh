#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
nproc; free -g | head -2
( time python bench.py --steps 20 --warmup 5 ) 2>&1 | tee gpurun_out/bench4.log | tail -5
( time python bench.py --impl reference --steps 3 --warmup 1 ) 2>&1 | tee gpurun_out/bench4_ref.log | tail -3
