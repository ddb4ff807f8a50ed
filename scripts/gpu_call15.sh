#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q 2>&1 | tail -2
for PF in 0 1 2 3; do
  echo "prefetch $PF"; python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --prefetch $PF 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('  ms/step', round(d['ms_per_step'],3), 'L0 kernel_ms', round(d['roofline']['kernel_ms'],3))"
done | tee gpurun_out/prefetch_ab.log
