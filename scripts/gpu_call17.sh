#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_surface.py -x -q --tb=short 2>&1 | tail -12 | tee gpurun_out/pytest17.log | tail -3
for FS in gather scatter gather scatter; do
  echo "fused-style $FS"; python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --fused-style $FS 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('  ms/step', round(d['ms_per_step'],3), 'GF', round(d['value']), 'L0 launch ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3))"
done | tee gpurun_out/fused_style_ab.log
