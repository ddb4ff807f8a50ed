"""Per-phase timing of the sharded step: torchrun --nproc-per-node N scripts/p2p_phase_probe.py [--blocks B]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
ap = argparse.ArgumentParser(); ap.add_argument("--blocks", type=int, default=1000); ap.add_argument("--k", type=int, default=128)
ap.add_argument("--exchange", default="p2p"); ap.add_argument("--overlap", type=int, default=0)
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from arrow_matrix_b200 import synth
from arrow_matrix_b200.comm import world_comm
from arrow_matrix_b200.sharded import ShardedArrowDecomposition
dec = synth.synth_decomposition(a.blocks, 10000, levels=2, perm_kind="random", seed=503)
arrow = ShardedArrowDecomposition(world_comm(), dec, 10000, a.k, device=rank, exchange=a.exchange, overlap=bool(a.overlap))
eng, be, ctx = arrow.engine, arrow.engine.be, arrow.engine.be.ctx
X = synth.generate_dense_matrix(eng.local_rows, a.k, np.float32, np.random.default_rng(rank))
eng.set_features(X)
def timed(name, fn, n=5):
    for _ in range(2): fn()
    ctx.sync(); dist.barrier()
    ctx.timer_start(3)
    for _ in range(n): fn()
    ctx.timer_stop(3)
    ms = ctx.timer_ms(3) / n
    t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(json.dumps({"phase": name, "ms_max_over_ranks": round(float(t.item()), 3), "ms_rank0": round(ms, 3)}), flush=True)
timed("barrier", lambda: be.barrier(), 20)
sh1, sh0 = eng.plan.levels[1], eng.plan.levels[0]
def fwd():
    be.pull_rows(dst=(1, eng.ci[1]), dst_off=sh1.hoff, src=(0, eng.xi[0]), src_bounds=sh0.bounds, row_map=eng.fwd[1], accumulate=False, forward=True); be.barrier()
timed("fwd_exchange+barrier", fwd)
timed("bcast_head_x2", lambda: [be.bcast_head((j, eng.xi[j]), 10000) for j in range(2)])
timed("spmm_L0", lambda: be.spmm(eng.mats[0], eng.tiles[0][0], eng.tiles[0][1]))
timed("spmm_L1", lambda: be.spmm(eng.mats[1], eng.tiles[1][0], eng.tiles[1][1]))
timed("reduce_head_x2", lambda: [be.reduce_head((j, 1), 10000) for j in range(2)])
def bwd():
    be.pull_rows(dst=(0, 1), dst_off=sh0.hoff, src=(1, 1), src_bounds=sh1.bounds, row_map=eng.bwd[0], accumulate=True, forward=False); be.barrier()
timed("bwd_exchange+barrier", bwd)
def full():
    eng.rewind_features(); eng.step()
timed("full_step", full)
timed("propagate_only", lambda: eng.propagate_features())
timed("aggregate_only", lambda: eng.aggregate())
dist.barrier(); dist.destroy_process_group()
