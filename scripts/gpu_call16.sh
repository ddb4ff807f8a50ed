#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q --tb=short 2>&1 | tail -12 | tee gpurun_out/pytest_multi2.log | tail -4
for OV in 1 0; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2p --overlap $OV --no-cpu 2>&1 | tee gpurun_out/bench16_n2_ov$OV.log | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('p2p overlap=$OV', 'ms/step', round(d['ms_per_step'],3), 'GF', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],1), d['e2e']['api'][:40], 'roof', d['roofline'] and round(d['roofline']['frac'],3))"
done
