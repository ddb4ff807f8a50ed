#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_surface.py -x -q --tb=short 2>&1 | tail -40 | tee gpurun_out/pytest_surface.log | grep -E "Error|error|assert|passed|failed" | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/p2p_phase_probe.py 2>&1 | tee gpurun_out/phase_probe_p2p.log | grep phase
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/p2p_phase_probe.py --exchange nccl 2>&1 | tee gpurun_out/phase_probe_nccl.log | grep phase
