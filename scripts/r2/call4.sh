#!/bin/bash
# round 2, GPU call 4 (1 GPU): rank-thread tests again (exchange tables now built at set-up), level-1 launch A/B, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ranks_one_gpu.py tests/test_gpu_15d.py tests/test_gpu_petsc.py -x -q --tb=short -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/c4_pytest.log | tail -12
timeout 300 python scripts/r2/level1_ab.py 1000 128 2>&1 | tee gpurun_out/c4_level1_ab.log
(cd ab_r01 && timeout 300 python - <<'PY' 2>&1 | tee ../gpurun_out/c4_level1_r01.log
import json, numpy as np
from arrow_matrix_b200 import synth
from arrow_matrix_b200.engine import ArrowEngine
dec = synth.synth_decomposition(1000, 10000, levels=2, perm_kind="random", seed=503)
eng = ArrowEngine(dec, 10000, 128, mode="fused"); ctx = eng.ctx
eng.set_features(synth.generate_dense_matrix(10000000, 128, np.float32, np.random.default_rng(1)))
st0, st1 = eng.levels; x = st0.bufs[st0.xi]
for _ in range(3): ctx.spmm(st1.csr_fused, x, st1.cbuf)
ctx.timer_start(3)
for _ in range(10): ctx.spmm(st1.csr_fused, x, st1.cbuf)
ctx.timer_stop(3)
print(json.dumps({"tree": "r01", "level1_ms": round(ctx.timer_ms(3) / 10, 4)}))
PY
)
# ncu: launch list of the default bench command + full capture of the two launches of one step (k = 128) and the k = 16 step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-verify > gpurun_out/c4_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_tiles -s 6 -c 2 -o gpurun_out/c4_prof_bench_k128 -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-verify --no-k16 > gpurun_out/c4_ncu_k128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_tiles -s 6 -c 2 -o gpurun_out/c4_prof_bench_k16 -f \
    python bench.py --k 16 --steps 2 --warmup 1 --no-e2e --no-cpu --no-verify > gpurun_out/c4_ncu_k16.log 2>&1
ls -la gpurun_out/c4_*
