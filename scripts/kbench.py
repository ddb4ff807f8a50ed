"""Kernel microbenchmark: one arrow block x k sweep x kernel variants.

Run on a B200 through gpurun.  Each timed launch is preceded by an L2 flush (256 MiB write) and
timed on the context's stream with CUDA events; prints one JSON line per (k, variant).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_matrix_b200 import _lib, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=100)
    ap.add_argument("--width", type=int, default=10000)
    ap.add_argument("--ks", type=str, default="16,32,64,128,256")
    ap.add_argument("--variants", type=str, default="0,1,3,19,35,67")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--uniform", action="store_true", help="utils.generate_sparse_matrix recipe instead of arrow shaped")
    ap.add_argument("--peak", type=float, default=0.0)
    ap.add_argument("--big-tiles", type=int, default=1)
    ap.add_argument("--prefetch", type=str, default="0", help="comma list of ARROW_OPT_PREFETCH values to sweep (low nibble = plain launches)")
    ap.add_argument("--carveout", type=str, default="-1", help="comma list of ARROW_OPT_SMEM_CARVEOUT values (percent; -1 = driver default)")
    ap.add_argument("--no-flush", action="store_true", help="skip the flushed timing (inputs far larger than L2)")
    args = ap.parse_args()

    peak = args.peak
    if not peak:
        try:
            peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            peak = 6650.0
    n = args.blocks * args.width
    rng = np.random.default_rng(42)
    t0 = time.time()
    if args.uniform:
        A = synth.generate_sparse_matrix(n, n, 10 * n, np.float32, rng)
    else:
        A = synth.arrow_csr(n, args.width, args.blocks, rng)
    print(f"# generated {n}x{n} nnz={A.nnz} in {time.time() - t0:.1f}s", flush=True)
    ctx = _lib.Context(0)
    ctx.set_option(ctx.OPT_BIG_TILES, args.big_tiles)
    dA = ctx.csr_from_scipy(A)
    # device-to-device copy bandwidth on this box (the roofline denominator's cross-check)
    a = ctx.dense_alloc(1 << 22, 64)
    b = ctx.dense_alloc(1 << 22, 64)
    for _ in range(3):
        b.copy_from(a)
    ctx.timer_start(0)
    for _ in range(10):
        b.copy_from(a)
    ctx.timer_stop(0)
    ms = ctx.timer_ms(0) / 10
    print(json.dumps({"probe": "d2d_copy", "GBps": 2 * (1 << 22) * 64 * 4 / ms / 1e6}), flush=True)
    a.free(); b.free()

    for k in [int(x) for x in args.ks.split(",")]:
        X = synth.generate_dense_matrix(n, k, np.float32, rng)
        dX, dC = ctx.dense_from_host(X), ctx.dense_alloc(n, k)
        alg_bytes = A.nnz * 8 + (n + 1) * 4 + 2.0 * n * k * 4
        flops = 2.0 * A.nnz * k
        for v, pf, cv in [(int(x), int(q), int(c)) for x in args.variants.split(",") for q in args.prefetch.split(",")
                          for c in args.carveout.split(",")]:
            if (v & 0xF) == 2 and (k < 32 or k > 128):
                continue
            if (v >> 8) == 2 and k > 32:
                continue
            ctx.set_option(ctx.OPT_PREFETCH, pf)
            ctx.set_option(ctx.OPT_SMEM_CARVEOUT, cv)
            for _ in range(3):
                ctx.spmm(dA, dX, dC, variant=v)
            times = []
            for it in range(args.iters):
                if not args.no_flush:
                    ctx.l2_flush()
                ctx.timer_start(1)
                ctx.spmm(dA, dX, dC, variant=v)
                ctx.timer_stop(1)
                times.append(ctx.timer_ms(1))
            # warm (no flush) timing too
            ctx.timer_start(2)
            for it in range(args.iters):
                ctx.spmm(dA, dX, dC, variant=v)
            ctx.timer_stop(2)
            warm = ctx.timer_ms(2) / args.iters
            med = float(np.median(times))
            print(json.dumps({"k": k, "variant": v, "rpg": (v >> 8) & 3, "vpl": (v >> 4) & 15, "prefetch": pf, "carveout": cv, "big_tiles": args.big_tiles, "rows": n, "ms_flushed": round(med, 4), "ms_min": round(min(times), 4),
                              "ms_back_to_back": round(warm, 4), "alg_GBps": round(alg_bytes / med / 1e6, 1),
                              "frac_of_peak": round(alg_bytes / med / 1e6 / peak, 3), "GFLOPs": round(flops / med / 1e6, 1),
                              "alg_MB": round(alg_bytes / 1e6, 1)}), flush=True)
        dX.free(); dC.free()
    ctx.close()


if __name__ == "__main__":
    main()
