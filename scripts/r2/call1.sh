#!/bin/bash
# round 2, GPU call 1 (1 GPU): regression tests of the rewritten tile kernel, gather-source probes, narrow-k variants,
# bulk L2 prefetch A/B, ncu captures of the k=16 level-0 launch and of the fused level-1 launch.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader > gpurun_out/c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15 | tee gpurun_out/c1_pytest.log | tail -4
timeout 120 python scripts/probe_gather.py 2>&1 | tee gpurun_out/c1_probe.log
# 1M-row level: rows-per-group 1 vs 2, prefetch none / current tile / look-ahead
timeout 300 python scripts/kbench.py --blocks 100 --iters 10 --ks 16,32 --variants 259,515 --prefetch 0,1,2 2>&1 | grep -v "^#" | tee gpurun_out/c1_kbench_1m_small.log | cut -c1-200
timeout 300 python scripts/kbench.py --blocks 100 --iters 10 --ks 64,128,256 --variants 3 --prefetch 0,1,2 2>&1 | grep -v "^#" | tee gpurun_out/c1_kbench_1m_big.log | cut -c1-200
# 10M-row level (bench size)
timeout 400 python scripts/kbench.py --blocks 1000 --iters 5 --no-flush --ks 16,32 --variants 259,515 --prefetch 0,2 2>&1 | grep -v "^#" | tee gpurun_out/c1_kbench_10m_small.log | cut -c1-200
timeout 400 python scripts/kbench.py --blocks 1000 --iters 5 --no-flush --ks 128 --variants 3 --prefetch 0,1,2 2>&1 | grep -v "^#" | tee gpurun_out/c1_kbench_10m_k128.log | cut -c1-200
# fused step: prefetch of the level-1 launch (high nibble) and of both
for pf in 0 16 32 34; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-verify --prefetch $pf 2>gpurun_out/c1_bench_pf$pf.err | tail -1 > gpurun_out/c1_bench_pf$pf.json
  python - <<PY
import json
d = json.load(open("gpurun_out/c1_bench_pf$pf.json"))
print("prefetch $pf: step", round(d["ms_per_step"], 3), "ms, level-0 launch", round(d["roofline"]["kernel_ms"], 3), "ms, frac", round(d["roofline"]["frac"], 3))
PY
done
