#!/bin/bash
mkdir -p gpurun_out
run() { # N exchange overlap tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29521 \
     bench.py --gpus $1 --steps 20 --warmup 5 --exchange $2 --overlap $3 --no-cpu 2>&1 | tee gpurun_out/bench19_$4.log | tail -1 | python -c "
import sys, json
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
    print('$4', 'ms/step', round(d['ms_per_step'],3), 'GF', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'e2e GF', round(d['e2e']['value']), 'roof', r.get('frac') and round(r['frac'],3), 'launches', d['gpu_launches'])
except Exception as e:
    print('$4 FAILED', e)"
}
run 8 p2p 1 n8_p2p_ov1
run 8 p2p 0 n8_p2p_ov0
run 4 p2p 1 n4_p2p_ov1
