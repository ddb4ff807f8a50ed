#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_surface.py tests/test_gpu_multi.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_2gpu.log | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/p2p_probe.py 2>&1 | tee gpurun_out/p2p_probe.log | tail -3
for H in 3,0 0,0 3,3 3,1 3,2 1,0; do
  echo "hints $H"; python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --l2-hints $H 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('  ms/step', round(d['ms_per_step'],3), 'L0 kernel_ms', round(d['roofline']['kernel_ms'],3))"
done | tee gpurun_out/hints_ab.log
