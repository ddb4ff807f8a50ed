"""GPU tests of the PETSc-style 1D baseline through the C ABI.

One-GPU cases run everywhere; the N-GPU halo exchange over NVLink peer memory is validated over gloo on CPU
(tests/test_petsc_baseline_cpu.py) and runs on hardware whenever the box has two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from arrow_matrix_b200 import synth
from arrow_matrix_b200.baseline import spmm_petsc
from arrow_matrix_b200.comm import SelfComm
from arrow_matrix_b200.matrix_slice import MatrixSlice
from tests.test_gpu_kernels import assert_close


def test_single_gpu_against_reference_golden(cuda_device):
    g = np.load(os.path.join(ROOT, "tests", "golden", "petsc_single_w1.npz"))
    n = int(g["n"])
    A = sparse.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=(n, n))
    sl = MatrixSlice.initialize(SelfComm(), A)
    eng = spmm_petsc.HaloSpmm(SelfComm(), sl, g["X"].shape[1], device=cuda_device)
    eng.set_features(g["X"])
    eng.spmm()
    assert_close(eng.result(), g["r0_Y"])
    eng.close()


@pytest.mark.parametrize("k", [4, 32, 128])
def test_spmm_gpu_reference_signature(cuda_device, k):
    """host arrays in, ``Y_i_local +=`` out, like spmm_petsc.py:229-326 (test_spmmPETSc.py:11-43 on one rank)"""
    rng = np.random.default_rng(1)
    n = 5000
    A = synth.generate_sparse_matrix(n, n, n * 9, np.float32, rng)
    sl = MatrixSlice.initialize(SelfComm(), A)
    X = synth.generate_dense_matrix(n, k, np.float32, rng)
    Y = np.ones((n, k), np.float32)
    out = spmm_petsc.spmm_gpu(SelfComm(), sl, X, Y, np.zeros((0, k), np.float32), device=cuda_device)
    assert out is Y
    assert_close(Y, 1.0 + (A.astype(np.float64) @ X.astype(np.float64)).astype(np.float32))
    X2 = synth.generate_dense_matrix(n, k, np.float32, rng)            # cached device state, fresh features
    Y2 = np.zeros((n, k), np.float32)
    spmm_petsc.spmm_gpu(SelfComm(), sl, X2, Y2, None, device=cuda_device)
    assert_close(Y2, (A.astype(np.float64) @ X2.astype(np.float64)).astype(np.float32))
    sl._halo_engine.close()


def test_benchmark_driver_synthetic(cuda_device, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    out = spmm_petsc.benchmark_spmm(None, 16, 3, "gpu", None, np.float32, np.random.default_rng(42), comm=SelfComm(),
                                    scale=512, verbose=False)
    assert len(out["times"]) == 3 and out["engine"].nnz > 0
    assert len(os.listdir(tmp_path / "logs")) == 4                    # the reference's four log artefacts
    with pytest.raises(NotImplementedError):
        spmm_petsc.benchmark_spmm(None, 16, 1, "cpu", None, np.float32, np.random.default_rng(42), comm=SelfComm(),
                                  scale=64, verbose=False)
    out["engine"].close()


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
        from arrow_matrix_b200.comm import world_comm
        sizes = np.array([3000 + 500 * r for r in range(world)], dtype=np.int64)
        n = int(sizes.sum())
        A = synth.generate_sparse_matrix(n, n, n * 8, np.float32, np.random.default_rng(9))
        bounds = np.concatenate([[0], np.cumsum(sizes)])
        s, e = int(bounds[rank]), int(bounds[rank + 1])
        comm = world_comm()
        sl = MatrixSlice.initialize(comm, A[s:e])
        eng = spmm_petsc.HaloSpmm(comm, sl, 64, device=rank, overlap=overlap)
        rng = np.random.default_rng(4)
        for _ in range(3):
            X = synth.generate_dense_matrix(n, 64, np.float32, rng)
            eng.set_features(X[s:e])
            eng.spmm()
            ref = (A.astype(np.float64) @ X.astype(np.float64)).astype(np.float32)[s:e]
            got = eng.result()
            scale = float(np.max(np.abs(ref)))
            assert float(np.max(np.abs(got - ref))) <= 1e-5 * scale
            assert np.array_equal(eng.halo(), X[sl.x_index_in])
        eng.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException:     # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("overlap", [True, False])
def test_halo_exchange_on_gpus(overlap):
    import torch.multiprocessing as mp
    world = min(_n_gpus(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, overlap, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    bad = [f"rank {r}: {m}" for r, m in sorted(results) if m != "ok"]
    assert not bad, "\n".join(bad)
