#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_full.log | tail -5
for BT in 1 0; do python scripts/kbench.py --blocks 100 --iters 10 --ks 16,32 --variants 3 --big-tiles $BT 2>&1 | grep '"k"' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('big_tiles=$BT k', d['k'], 'ms', d['ms_flushed'], 'frac', d['frac_of_peak'])"; done | tee gpurun_out/kbench13.log
for OV in 1 0; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2p --overlap $OV --no-cpu 2>&1 | tee gpurun_out/bench13_n2_ov$OV.log | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('p2p overlap=$OV', 'ms/step', round(d['ms_per_step'],3), 'GF', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],1), d['e2e']['api'][:40])"
done
