"""Multi-GPU tuning sweep (torchrun): one engine per process group, many schedules.

Prints one JSON line per configuration: step time (max over ranks) of the fused step for lane / CTA-cap / push-grid /
graph-replay settings, the 'dry' step (no cross-GPU effects), per-kernel times, and the exchange-mode step for reference."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    a = bench.parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    from arrow_matrix_b200 import _lib, graphio, comm as comm_mod
    _lib.bind_thread_to_device_numa(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    comm = comm_mod.world_comm()
    tag = f"{a.blocks}_{a.perm}"
    base = os.path.join(ROOT, "tmp", f"bench_{tag}_{a.width}_{a.levels}")
    if rank == 0 and not os.path.exists(base + f"_B_{a.width}_0_bd_indptr.npy"):
        graphio.save_decomposition_new(bench.build_decomposition(a), base, a.width, block_diagonal=True)
    comm.Barrier()

    def mx(x):
        return bench.max_over_ranks(dist, torch, x)

    def out(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    lean = os.environ.get("SWEEP_LEAN") == "1"
    for mode in (("auto",) if lean else ("auto", "exchange")):
        a.mode = mode
        arrow, eng, blocks = bench.build_engine(a, comm, base, a.k, local_rank)
        ctx = eng.ctx
        rows_local = eng.local_rows
        X = _lib.PinnedArray((rows_local, a.k))
        X.array[:] = 2 * np.random.default_rng(42 + rank).random((rows_local, a.k), dtype=np.float32) - 1
        eng.set_features(X.array)
        ctx.sync()

        def barrier():
            dist.barrier()
            eng.sync()

        if eng.fp is None:
            for ov in (1, 0):
                eng.overlap = bool(ov)
                ms = mx(bench.time_steps(eng, ctx, barrier, a.steps, 3))
                out({"n": world, "k": a.k, "mode": eng.mode, "overlap": ov, "ms_per_step": round(ms, 4)})
            eng.close()
            continue
        configs = []
        for ov, mc, sc in ((1, 2, 2), (1, 3, 1), (1, 1, 3), (1, 3, 2), (1, 2, 3), (1, 4, 4), (0, 0, 0)):
            configs.append((1, ov, mc, sc, 0))
        configs.append((0, 1, 2, 2, 0))
        for pc in (74, 148):
            configs.append((1, 1, 2, 2, pc))
        if lean:
            configs = [(1, 1, 2, 2, 0), (1, 1, 3, 2, 0), (1, 0, 0, 0, 0)]
        for bwd, il in (("push", 1), ("pull", 1), ("pull", 0), ("push", 0)):
            eng.bwd_mode = bwd
            ctx.set_option(ctx.OPT_PUSH_INTERLEAVE, il)
            eng.overlap, eng.main_ctas, eng.side_ctas, eng.use_graphs = True, 2, 2, True
            eng.graphs, eng._graph_warm = {}, set()
            ms = mx(bench.time_steps(eng, ctx, barrier, a.steps, 4))
            out({"n": world, "k": a.k, "mode": eng.mode, "bwd_mode": bwd, "push_interleave": il, "ms_per_step": round(ms, 4)})
        eng.bwd_mode = "push"
        ctx.set_option(ctx.OPT_PUSH_INTERLEAVE, 1)
        for graphs, ov, mc, sc, pc in configs:
            eng.overlap = bool(ov)
            eng.main_ctas, eng.side_ctas = mc, sc
            ctx.set_option(ctx.OPT_PUSH_CTAS, pc)
            eng.use_graphs = bool(graphs)
            eng.graphs, eng._graph_warm = {}, set()
            ms = mx(bench.time_steps(eng, ctx, barrier, a.steps, 4))
            out({"n": world, "k": a.k, "mode": eng.mode, "graphs": graphs, "overlap": ov, "main_ctas": mc, "side_ctas": sc,
                 "push_ctas": pc, "ms_per_step": round(ms, 4)})
        ctx.set_option(ctx.OPT_PUSH_CTAS, 0)
        eng.use_graphs = False
        eng.overlap, eng.main_ctas, eng.side_ctas = True, 2, 2
        dry = mx(bench.time_steps(eng, ctx, barrier, a.steps, 3, step_fn=lambda: eng._step_fused(dry=True)))
        eng.overlap = False
        dry0 = mx(bench.time_steps(eng, ctx, barrier, a.steps, 3, step_fn=lambda: eng._step_fused(dry=True)))
        out({"n": world, "k": a.k, "dry_ms_overlap": round(dry, 4), "dry_ms_serial": round(dry0, 4)})
        # per-kernel times on this rank's shard (max over ranks), everything on the main lane
        be = eng.be
        x = (0, eng.xi[0])

        def t(fn, iters=10):
            for _ in range(2):
                fn()
            barrier()
            ctx.timer_start(4)
            for _ in range(iters):
                fn()
            ctx.timer_stop(4)
            ms = ctx.timer_ms(4) / iters
            barrier()
            return mx(ms)
        L = eng.L
        phases = {"push": t(lambda: be.push(eng.f_push, x)),
                  "barrier": t(lambda: be.barrier()),
                  "l0_spmm": t(lambda: be.spmm(eng.mats[0], eng.tiles[0][x[1]], eng.tiles[0][x[1] ^ 1])),
                  "final_add": t(lambda: be.final_add((0, x[1] ^ 1), eng.plan.levels[0].hoff, eng.plan.levels[0].own_rows, eng._stg[0], eng.f_add[0]))}
        for j in range(1, L):
            if eng.f_mats[j] is not None:
                phases[f"l{j}_spmm"] = t(lambda j=j: be.spmm_fused(eng.f_mats[j], x, eng._recv, eng.fp.x_split, eng.f_tables[j]))

            def copies(j=j):
                for sr, src_off, rows, dst_off in eng.fp.recv_plan[j]:
                    be.copy_rows_from_peer(dst=eng._stg[j - 1], dst_off=dst_off, peer=sr, src=eng._snd[j], src_off=src_off, rows=rows)
            phases[f"l{j}_pulls"] = t(copies)
            phases[f"l{j}_bpush"] = t(lambda j=j: be.push(eng.f_bpush[j], eng._snd[j]))
        fp = eng.fp
        out({"n": world, "k": a.k, "phases_ms": {kk: round(v, 4) for kk, v in phases.items()},
             "recv_rows_rank0": int(fp.recv_rows), "push_rows_rank0": int(fp.push_bounds[-1]), "stage_rows_rank0": fp.stage_rows,
             "partition": eng.plan.partition_used})
        eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
