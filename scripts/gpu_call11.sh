#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_gpu_surface.py -x -q --tb=short 2>&1 | tail -30 | tee gpurun_out/pytest_multi.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/p2p_phase_probe.py 2>&1 | tee gpurun_out/phase_probe_p2p_packed.log | grep phase
for CFG in "p2p 1" "p2p 0" "nccl 0"; do
  set -- $CFG
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 5 --exchange $1 --overlap $2 --no-cpu 2>&1 | tee gpurun_out/bench11_n2_$1_ov$2.log | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$1 overlap=$2', 'ms/step', round(d['ms_per_step'],3), 'GF', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'launches', d['gpu_launches'])"
done
