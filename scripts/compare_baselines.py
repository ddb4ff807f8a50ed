#!/usr/bin/env python
"""Arrow decomposition vs the 1D (PETSc-style) and 1.5D baselines on the SAME graph, same kernels, N GPUs.

    torchrun --nproc-per-node N scripts/compare_baselines.py --vertices 2000000 --neighbors 8 --width 20000 -k 128

The paper's comparison (reference: arrow_bench.py vs baseline/spmm_petsc.py vs baseline/spmm_15d.py) re-run on B200:
one Barabasi-Albert graph, decomposed with ``arrow_decomposition`` for the arrow engine and cut into row slices for
the baselines.  Every rank prints nothing; rank 0 prints one JSON line per algorithm with device-timed ms per product
(CUDA events on the launching stream, max over ranks) and GFLOP/s.  NOT part of bench.py's contract; prepared for the
round-2 measurements (the N-GPU paths of the two baselines are gloo-validated only so far).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(ctx, comm, fn, steps, warmup):
    """device time of ``steps`` calls of ``fn`` (ms per call, max over ranks)"""
    for _ in range(warmup):
        fn()
    ctx.sync()
    comm.Barrier()
    ctx.timer_start(7)
    for _ in range(steps):
        fn()
    ctx.timer_stop(7)
    ms = ctx.timer_ms(7) / steps
    return max(comm.allgather(float(ms)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vertices", type=int, default=1000000)
    ap.add_argument("--neighbors", type=int, default=8)
    ap.add_argument("--width", type=int, default=10000)
    ap.add_argument("-k", type=int, default=128)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--replication", type=int, default=0)
    ap.add_argument("--skip", type=str, default="", help="comma list of: arrow,petsc,15d")
    a = ap.parse_args()

    from arrow_matrix_b200 import comm as comm_mod, synth
    from arrow_matrix_b200.baseline import spmm_15d, spmm_petsc
    from arrow_matrix_b200.decomposition import arrow_decomposition
    from arrow_matrix_b200.matrix_slice import MatrixSlice
    from arrow_matrix_b200.sharded import ShardedArrowDecomposition

    comm_mod.init_from_env()
    comm = comm_mod.world_comm()
    rank, world = comm.Get_rank(), comm.Get_size()
    device = int(os.environ.get("LOCAL_RANK", "0"))
    skip = set(a.skip.split(",")) if a.skip else set()
    A = synth.barabasi_albert(a.vertices, a.neighbors, 503)            # same seed on every rank: same graph
    n, k = A.shape[0], a.k
    rng = np.random.default_rng(42)
    out = []

    def report(name, ms, extra, nnz_multiplied=None):
        if rank == 0:
            nnz_m = int(A.nnz if nnz_multiplied is None else nnz_multiplied)
            line = dict(algorithm=name, n_gpus=world, vertices=n, nnz=int(A.nnz), nnz_multiplied=nnz_m, k=k,
                        ms_per_product=ms, gflops=2.0 * nnz_m * k / ms / 1e6, **extra)
            print(json.dumps(line), flush=True)
            out.append(line)

    if "arrow" not in skip:
        dec = arrow_decomposition(A, a.width, a.levels, block_diagonal=True, seed=503)
        if world == 1:
            from arrow_matrix_b200.engine import ArrowEngine
            eng = ArrowEngine(dec, a.width, k, device=device)
            eng.set_features(synth.generate_dense_matrix(eng.levels[0].rows, k, np.float32, rng))
            ms = timed(eng.ctx, comm, eng.step, a.steps, a.warmup)
            # entries of the best-effort last level outside the arrow pattern are dropped, like in the reference
            report("arrow", ms, dict(levels=len(dec), mode=eng.mode), nnz_multiplied=eng.total_nnz)
            eng.close()
        else:
            arrow = ShardedArrowDecomposition(comm, dec, a.width, k, device=device, exchange="p2p", overlap=True)
            sh0 = arrow.engine.plan.levels[0]
            arrow.set_features(synth.generate_dense_matrix(sh0.own_rows, k, np.float32, rng))
            ms = timed(arrow.engine.ctx, comm, arrow.step, a.steps, a.warmup)
            report("arrow", ms, dict(levels=len(dec)), nnz_multiplied=arrow.engine.total_nnz)

    bounds = (np.arange(world + 1, dtype=np.int64) * n + world - 1) // world
    if "petsc" not in skip:
        s, e = int(bounds[rank]), int(bounds[rank + 1])
        sl = MatrixSlice.initialize(comm, A[s:e])
        eng = spmm_petsc.HaloSpmm(comm, sl, k, device=device)
        eng.set_features(synth.generate_dense_matrix(e - s, k, np.float32, rng))
        ms = timed(eng.fab.ctx, comm, eng.spmm, a.steps, a.warmup)
        halo = comm.allgather(int(eng.n_halo))
        report("petsc_1d", ms, dict(halo_rows_max=max(halo), halo_rows_total=sum(halo)))
        eng.close()

    if "15d" not in skip:
        c = a.replication or spmm_15d.largest_power_of_two_square(world)
        lA, lX, lY, grid, _, _, _ = spmm_15d.generate_15d_decomposition(A, k, np.float32, c, rng, comm=comm)
        eng = spmm_15d.Spmm15D(grid, lA, lX.shape[0], k, device=device)
        eng.set_features(lX)
        ms = timed(eng.fab.ctx, comm, eng.spmm, a.steps, a.warmup)
        report("spmm_15d", ms, dict(replication=c, rounds=grid.rounds))
        eng.close()
    comm.Barrier()
    return out


if __name__ == "__main__":
    main()
