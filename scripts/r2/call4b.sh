#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_ranks_one_gpu.py -x -q --tb=short -p no:cacheprovider > gpurun_out/c4b_pytest_full.log 2>&1
head -60 gpurun_out/c4b_pytest_full.log | cut -c1-220; echo ...; tail -8 gpurun_out/c4b_pytest_full.log | cut -c1-220
timeout 300 python scripts/r2/level1_ab.py 1000 128 2>&1 | tee gpurun_out/c4b_level1_ab.log
timeout 300 python scripts/r2/level1_ab.py 1000 16 2>&1 | tee gpurun_out/c4b_level1_ab_k16.log
