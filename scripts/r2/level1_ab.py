"""A/B of the fused level-1 launch (C_1 = B_1 X0[cmap], scattered first-touch rows): unpredicated batches vs the predicated
gather path (ARROW_OPT_FORCE_PREDICATED), and the level-0 launch for reference.  One JSON line per variant."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from arrow_matrix_b200 import synth  # noqa: E402
from arrow_matrix_b200.engine import ArrowEngine  # noqa: E402

blocks, w, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 128
dec = synth.synth_decomposition(blocks, w, levels=2, perm_kind="random", seed=503)
eng = ArrowEngine(dec, w, k, mode="fused")
ctx = eng.ctx
X = synth.generate_dense_matrix(blocks * w, k, np.float32, np.random.default_rng(1))
eng.set_features(X)
st0, st1 = eng.levels
x = st0.bufs[st0.xi]
out = st0.bufs[1 - st0.xi]


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    ctx.timer_start(3)
    for _ in range(iters):
        fn()
    ctx.timer_stop(3)
    return ctx.timer_ms(3) / iters


for kernel in (1, 0):
    ctx.set_option(ctx.OPT_TILE_KERNEL, kernel)
    for forced in (0, 1):
        ctx.set_option(ctx.OPT_FORCE_PREDICATED, forced)
        ms1 = timed(lambda: ctx.spmm(st1.csr_fused, x, st1.cbuf))
        ms0 = timed(lambda: ctx.spmm_add(st0.csr, x, out, st1.cbuf, st1.to_next_dev))
        eng.rewind_features()
        ms = timed(eng.step)
        print(json.dumps({"k": k, "tile_kernel": "round-1" if kernel else "generalised", "predicated": forced, "level1_ms": round(ms1, 4),
                          "level0_add_ms": round(ms0, 4), "step_ms": round(ms, 4)}), flush=True)
ctx.set_option(ctx.OPT_FORCE_PREDICATED, 0)
ctx.set_option(ctx.OPT_TILE_KERNEL, 1)
