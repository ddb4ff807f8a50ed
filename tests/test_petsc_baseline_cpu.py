"""CPU tests of the PETSc-style 1D baseline: MatrixSlice tables against golden vectors produced by the
unmodified reference (tests/golden/make_golden_petsc.py), and the halo-exchange engine's host logic over gloo with
numpy tiles, with the assertions of the reference's own tests (tests/test_spmmPETSc.py:36-43)."""
import glob
import os
import socket
import sys

import numpy as np
import pytest
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "petsc_*.npz")))
TABLES = ["x_index_in", "rank_in", "x_index_out", "rank_out", "send_count", "recv_count", "x_index_out_localized",
          "x_index_in_localized", "send_sdispl", "recv_sdispl"]


class _ListComm:
    """Rank ``r`` of a world whose allgather / alltoall answers are computed from per-rank inputs known up front
    (single process): enough for MatrixSlice.initialize, which only exchanges n_i and the halo index lists."""

    def __init__(self, rank, world, n_i_all, wanted_matrix=None):
        self.r, self.w, self.n_i_all, self.wanted = rank, world, n_i_all, wanted_matrix

    def Get_rank(self):
        return self.r

    def Get_size(self):
        return self.w

    def Barrier(self):
        pass

    def allgather(self, obj):
        assert isinstance(obj, int)
        return list(self.n_i_all)

    def alltoall(self, objs):
        return [self.wanted[s][self.r] for s in range(self.w)]


def _load(path):
    g = np.load(path, allow_pickle=False)
    n = int(g["n"])
    A = sparse.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=(n, n))
    return g, A, g["X"], g["all_n_i"].astype(np.int64)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[6:-4] for p in GOLD])
def test_matrix_slice_tables_match_reference(path):
    from arrow_matrix_b200.matrix_slice import MatrixSlice
    g, A, X, all_n_i = _load(path)
    world = int(g["world"])
    bounds = np.concatenate([[0], np.cumsum(all_n_i)])
    # what every rank asks every other rank for (the all-to-all payload), computed per rank first
    wanted = []
    for r in range(world):
        cols, x_in, rank_in = MatrixSlice.construct_receive_tables(A[bounds[r]:bounds[r + 1]], bounds[r], bounds[r + 1], all_n_i)
        wanted.append([x_in[rank_in == d] for d in range(world)])
    for r in range(world):
        sl = MatrixSlice.initialize(_ListComm(r, world, all_n_i, wanted), A[bounds[r]:bounds[r + 1]])
        for key in TABLES:
            assert np.array_equal(np.asarray(getattr(sl, key)), g[f"r{r}_{key}"]), (r, key)
        assert sl.start_col == int(g[f"r{r}_start_col"]) and sl.end_col == int(g[f"r{r}_end_col"])
        for name, M in (("loc", sl.A_i_local), ("non", sl.A_i_nonlocal)):
            M = sparse.csr_matrix(M)
            M.sort_indices()
            assert tuple(M.shape) == tuple(g[f"r{r}_{name}_shape"])
            ref = sparse.csr_matrix((g[f"r{r}_{name}_data"], g[f"r{r}_{name}_indices"], g[f"r{r}_{name}_indptr"]), shape=M.shape)
            assert abs(M - ref).nnz == 0
        # the reference's result from these pieces
        Xn = X[sl.x_index_in] if sl.x_index_in.size else np.zeros((0, X.shape[1]), X.dtype)
        Y = sl.A_i_local @ X[bounds[r]:bounds[r + 1]] + sl.A_i_nonlocal @ Xn
        assert np.allclose(Y, g[f"r{r}_Y"], rtol=1e-5, atol=1e-6)
        assert np.array_equal(Xn, g[f"r{r}_X_nonlocal"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cases, q):
    """one process per rank, all cases of a world size in one process group; a failing rank reports and exits"""
    current = None
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        for case, overlap in cases:
            current = (case, overlap)
            _run_case(rank, world, case, overlap)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException:     # noqa: BLE001
        import traceback
        q.put((rank, f"FAIL in case {current}: " + traceback.format_exc()))


def _run_case(rank, world, case, overlap):
    from arrow_matrix_b200.comm import world_comm
    from arrow_matrix_b200.matrix_slice import MatrixSlice
    from arrow_matrix_b200.baseline.spmm_petsc import HaloSpmm
    from tests.numpy_backend import GlooNumpyHaloFabric
    comm = world_comm()
    if case.startswith("golden:"):
        g, A, X, all_n_i = _load(os.path.join(ROOT, "tests", "golden", f"petsc_{case[7:]}.npz"))
        assert int(g["world"]) == world
        Xs = [X]
    else:
        rng = np.random.default_rng(3)
        sizes = {"unequal": [33] * (world // 2 + world % 2) + [5] * (world // 2), "zero": [40] + [0] * (world - 1),
                 "big": [700] * world}[case]
        all_n_i = np.array(sizes, dtype=np.int64)
        n = int(all_n_i.sum())
        A = sparse.random(n, n, density=0.03 if case != "big" else 0.004, format="csr", random_state=7, dtype=np.float32)
        Xs = [np.round(5 * rng.random((n, 8))).astype(np.float32), (2 * rng.random((n, 8)) - 1).astype(np.float32)]
        g = None
    bounds = np.concatenate([[0], np.cumsum(all_n_i)])
    s, e = int(bounds[rank]), int(bounds[rank + 1])
    sl = MatrixSlice.initialize(comm, A[s:e])
    assert MatrixSlice.check_comm_tables(comm, sl.x_index_in, sl.rank_in, sl.x_index_out, sl.rank_out)
    if g is not None:
        for key in TABLES:
            assert np.array_equal(np.asarray(getattr(sl, key)), g[f"r{rank}_{key}"]), key
    fab = GlooNumpyHaloFabric(comm)
    eng = HaloSpmm(comm, sl, Xs[0].shape[1], fabric=fab, overlap=overlap)
    assert eng.overlap == (overlap and world > 1)
    for X in Xs:                                           # fresh features per product, state reused
        eng.set_features(X[s:e])
        eng.spmm()
        Y = eng.result()
        ref = (A @ X)[s:e]
        assert Y.shape == ref.shape and np.allclose(Y, ref, rtol=1e-5, atol=1e-5), (case, rank)
        assert np.array_equal(eng.halo(), X[sl.x_index_in] if sl.x_index_in.size else np.zeros((0, X.shape[1]), np.float32))
        if g is not None:
            assert np.allclose(Y, g[f"r{rank}_Y"], rtol=1e-5, atol=1e-6)
    assert fab.n_barriers == (2 * len(Xs) if world > 1 else 0)


CASES = [(2, "golden:unequal_w2", True), (3, "golden:unequal_w3", True), (4, "golden:unequal_w4", False),
         (4, "golden:hubs_w4", True), (3, "golden:eye_w3", True), (2, "golden:empty_w2", False),
         (2, "unequal", False), (3, "zero", True), (3, "big", True)]


@pytest.mark.parametrize("world", [2, 3, 4])
def test_halo_engine_over_gloo(world):
    import torch.multiprocessing as mp
    cases = [(c, o) for w, c, o in CASES if w == world]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(30)
    bad = [f"rank {rank}: {msg}" for rank, msg in sorted(results) if msg != "ok" and "Connection closed by peer" not in msg]
    bad = bad or [f"rank {rank}: {msg}" for rank, msg in sorted(results) if msg != "ok"]
    assert not bad, "\n".join(bad)


def test_single_rank_engine_and_reference_facing_errors():
    sys.path.insert(0, ROOT)
    from arrow_matrix_b200.comm import SelfComm
    from arrow_matrix_b200.matrix_slice import MatrixSlice
    from arrow_matrix_b200.baseline import spmm_petsc
    from tests.numpy_backend import GlooNumpyHaloFabric
    g, A, X, _ = _load(os.path.join(ROOT, "tests", "golden", "petsc_single_w1.npz"))
    sl = MatrixSlice.initialize(SelfComm(), A)
    assert sl.x_index_in.size == 0 and sl.A_i_nonlocal.shape == (40, 0)
    eng = spmm_petsc.HaloSpmm(SelfComm(), sl, X.shape[1], fabric=GlooNumpyHaloFabric(SelfComm()))
    eng.set_features(X)
    eng.spmm()
    assert np.allclose(eng.result(), g["r0_Y"], rtol=1e-5, atol=1e-6)
    with pytest.raises(NotImplementedError):
        spmm_petsc.spmm_cpu(SelfComm(), sl, X, np.zeros_like(X), np.zeros((0, X.shape[1]), np.float32))
    with pytest.raises(ValueError):
        MatrixSlice.initialize(SelfComm(), A[:10])            # not square over the ranks (matrix_slice.py:118-121)
    with pytest.raises(ValueError):
        eng.set_features(X[:5])
