"""GPU tests of the 1.5D baseline.  The engine's host logic and its parity with the reference are
covered on CPU (tests/test_15d_baseline_cpu.py); here the same engine runs on hardware: one GPU against the reference's
golden run, and a P = 4 (or 2) grid when the box has the GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def test_single_gpu_against_reference_golden(cuda_device):
    from arrow_matrix_b200.baseline import spmm_15d
    from arrow_matrix_b200.comm import SelfComm
    from tests.test_gpu_kernels import assert_close
    g = np.load(os.path.join(ROOT, "tests", "golden", "spmm15d_p1_c1.npz"))
    n = int(g["n"])
    A = sparse.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=(n, n))
    lA, X, Y, grid, _, _, _ = spmm_15d.generate_15d_decomposition(A, int(g["k"]), np.float32, 1, None, comm=SelfComm(),
                                                                   X_full=g["X_full"])
    out = spmm_15d.spmm_15d_gpu(lA, X, Y, grid, device=cuda_device)
    assert out is Y
    assert_close(Y, g["r0_Y"])
    grid._engine.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, c, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
        from arrow_matrix_b200 import synth
        from arrow_matrix_b200.baseline import spmm_15d
        from arrow_matrix_b200.comm import world_comm
        n, k = 20011, 64
        A = synth.generate_sparse_matrix(n, n, n * 8, np.float32, np.random.default_rng(9))
        Xf = synth.generate_dense_matrix(n, k, np.float32, np.random.default_rng(4))
        lA, X, Y, grid, _, _, _ = spmm_15d.generate_15d_decomposition(A, k, np.float32, c, None, comm=world_comm(), X_full=Xf)
        ref = (A.astype(np.float64) @ Xf.astype(np.float64)).astype(np.float32)
        lNI = -(-n // grid.p_div_c)
        for _ in range(2):
            spmm_15d.spmm_15d_gpu(lA, X, Y, grid, device=rank)
            want = ref[grid.x * lNI:(grid.x + 1) * lNI]
            assert Y.shape == want.shape and float(np.max(np.abs(Y - want))) <= 1e-5 * float(np.max(np.abs(ref)))
        grid._engine.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException:     # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("c", [1, 2])
def test_15d_on_gpus(c):
    import torch.multiprocessing as mp
    world = 4 if _n_gpus() >= 4 else 2
    if world // c < c:
        pytest.skip("replication factor 2 needs a 4-rank grid")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, c, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    bad = [f"rank {r}: {m}" for r, m in sorted(results) if m != "ok"]
    assert not bad, "\n".join(bad)
