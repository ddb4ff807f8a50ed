#!/bin/bash
# one-GPU validation of the PETSc-style baseline on the C ABI + the whole GPU parity suite once more
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_call22.log | tail -12
