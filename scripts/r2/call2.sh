#!/bin/bash
# round 2, GPU call 2 (1 GPU): same-box A/B of the round-1 tree (ab_r01/) against the current one; L1 carve-out sweep;
# rows-per-group and tile size at k = 16 / 32.
mkdir -p gpurun_out
echo "== r01 tree"; (cd ab_r01 && timeout 300 python scripts/kbench.py --blocks 1000 --iters 5 --ks 16,32,128 --variants 3 2>&1 | grep -v "^#" | tee ../gpurun_out/c2_r01_kbench_10m.log | cut -c1-160)
echo "== current tree"
timeout 400 python scripts/kbench.py --blocks 1000 --iters 5 --no-flush --ks 16,32 --variants 259,515 --carveout=-1,25,50,100 2>&1 | grep -v "^#" | tee gpurun_out/c2_kbench_10m_small.log | cut -c1-230
timeout 400 python scripts/kbench.py --blocks 1000 --iters 5 --no-flush --ks 16,32 --variants 259 --big-tiles 0 --carveout=-1,25 2>&1 | grep -v "^#" | tee gpurun_out/c2_kbench_10m_small_t64.log | cut -c1-230
timeout 400 python scripts/kbench.py --blocks 1000 --iters 5 --no-flush --ks 128 --variants 3 --carveout=-1,25,50,100 2>&1 | grep -v "^#" | tee gpurun_out/c2_kbench_10m_k128.log | cut -c1-230
echo "== bench, r01 tree then current"
(cd ab_r01 && timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-verify 2>../gpurun_out/c2_bench_r01.err | tail -1 > ../gpurun_out/c2_bench_r01.json)
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-verify 2>gpurun_out/c2_bench_cur.err | tail -1 > gpurun_out/c2_bench_cur.json
python - <<PY
import json
for n in ("r01", "cur"):
    d = json.load(open(f"gpurun_out/c2_bench_{n}.json"))
    print(n, ": step", round(d["ms_per_step"], 3), "ms, level-0 launch", round(d["roofline"]["kernel_ms"], 3), "ms, frac", round(d["roofline"]["frac"], 3))
PY
