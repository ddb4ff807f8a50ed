"""2-rank probe of the NVLink peer path: torchrun --nproc-per-node 2 scripts/p2p_probe.py"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from arrow_matrix_b200 import _lib
from arrow_matrix_b200.comm import world_comm
comm = world_comm()
ctx = _lib.Context(rank)
rows, k = 2_000_000, 128                       # 1.024 GB tile
tile = ctx.dense_alloc(rows, k); tile.fill(float(rank + 1)); ctx.sync()
handles = comm.allgather(tile.ipc_export())
peer = ctx.ipc_import(handles[1 - rank], rows, k)
dst = ctx.dense_alloc(rows, k)
rng = np.random.default_rng(0)
ident = ctx.map_upload(np.arange(rows, dtype=np.int64), rows)
perm = ctx.map_upload(rng.permutation(rows).astype(np.int64), rows)
def t(fn, n=5):
    for _ in range(2): fn()
    ctx.timer_start(0)
    for _ in range(n): fn()
    ctx.timer_stop(0)
    return ctx.timer_ms(0) / n
gb = rows * k * 4 / 1e9
res = {}
res["memcpy_peer_GBps"] = gb / t(lambda: dst.copy_from(peer)) * 1e3
res["memcpy_local_GBps"] = gb / t(lambda: dst.copy_from(tile)) * 1e3
res["gather_ident_peer_GBps"] = gb / t(lambda: ctx.gather_rows(dst, peer, ident)) * 1e3
res["gather_perm_peer_GBps"] = gb / t(lambda: ctx.gather_rows(dst, peer, perm)) * 1e3
res["gather_ident_local_GBps"] = gb / t(lambda: ctx.gather_rows(dst, tile, ident)) * 1e3
res["gather_perm_local_GBps"] = gb / t(lambda: ctx.gather_rows(dst, tile, perm)) * 1e3
res["gatheradd_perm_peer_GBps"] = gb / t(lambda: ctx.gather_rows(dst, peer, perm, accumulate=True)) * 1e3
got = dst.d2h(rows=4)
dist.barrier()
if rank == 0:
    print(json.dumps({k_: round(v, 1) for k_, v in res.items()}), flush=True)
# NCCL reference: all_to_all of the same volume
x = torch.empty(rows * k // 2 * 2, device="cuda"); y = torch.empty_like(x)
for _ in range(2): dist.all_to_all_single(y, x)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5): dist.all_to_all_single(y, x)
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(json.dumps({"nccl_a2a_send_half_GBps": round(gb / 2 / (e0.elapsed_time(e1) / 5) * 1e3, 1)}), flush=True)
dist.destroy_process_group()
