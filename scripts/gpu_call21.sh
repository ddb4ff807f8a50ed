#!/bin/bash
# last sanity of the round: GPU parity suite + smoke after the C-ABI clean-up (no bench: budget)
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_call21.log | tail -12
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
