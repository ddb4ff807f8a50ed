#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_petsc.py tests/test_gpu_engine.py -m gpu -x -q --tb=short 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_call23.log | tail -12
