"""CPU tests of the 1.5D baseline: decomposition and product against golden vectors produced by the
unmodified reference on a thread MPI grid (tests/golden/make_golden_15d.py); engine host logic over gloo with numpy
tiles (snapshot-at-barrier peer reads)."""
import glob
import os
import socket
import sys

import numpy as np
import pytest
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "spmm15d_*.npz")))


def _load(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"spmm15d_{name}.npz"), allow_pickle=False)
    n = int(g["n"])
    A = sparse.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=(n, n))
    return g, A


def _check_rank(g, rank, lA, X, lNKb, grid):
    assert lNKb == int(g[f"r{rank}_lNKb"]) and len(lA) == int(g[f"r{rank}_n_blocks"])
    assert [grid.x, grid.y] == list(g[f"r{rank}_coords"])
    assert np.array_equal(X, g[f"r{rank}_X"])
    for r, B in enumerate(lA):
        assert tuple(B.shape) == tuple(g[f"r{rank}_A{r}_shape"]), (rank, r)
        ref = sparse.csr_matrix((g[f"r{rank}_A{r}_data"], g[f"r{rank}_A{r}_indices"], g[f"r{rank}_A{r}_indptr"]), shape=B.shape)
        assert abs(sparse.csr_matrix(B) - ref).nnz == 0


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
        for name, root_only in cases:
            current = name
            _run_case(rank, name, root_only)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException:     # noqa: BLE001
        import traceback
        q.put((rank, f"FAIL in case {current}: " + traceback.format_exc()))


def _run_case(rank, name, root_only):
    from arrow_matrix_b200.comm import world_comm
    from arrow_matrix_b200.baseline import spmm_15d
    from tests.numpy_backend import GlooNumpyHaloFabric
    g, A = _load(name)
    c, k = int(g["c"]), int(g["k"])
    comm = world_comm()
    src = A if (rank == 0 or not root_only) else None
    if not root_only and rank % 2 == 1:                     # triplet input like generate_15d_decomposition_new
        src = (A.data, A.indices, A.indptr)
    fn = spmm_15d.generate_15d_decomposition_new if isinstance(src, tuple) else spmm_15d.generate_15d_decomposition
    lA, X, Y, grid, _, _, lNKb = fn(src, k, np.float32, c, None, comm=comm, X_full=g["X_full"])
    _check_rank(g, rank, lA, X, lNKb, grid)
    fab = GlooNumpyHaloFabric(comm)
    eng = spmm_15d.Spmm15D(grid, lA, X.shape[0], k, fabric=fab)
    for it in range(2):                                     # state is reusable
        eng.set_features(X)
        eng.spmm()
        got = eng.result()
        assert got.shape == g[f"r{rank}_Y"].shape
        assert np.allclose(got, g[f"r{rank}_Y"], rtol=1e-5, atol=1e-6), (name, rank)
    assert fab.n_barriers == 2 * (3 if c > 1 else 2)


CASES = [("p2_c1", True), ("p4_c1", False), ("p4_c2", True), ("p4_c2_ragged", False), ("p8_c2", False)]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_15d_engine_over_gloo(world):
    import torch.multiprocessing as mp
    cases = [(n, r) for n, r in CASES if int(_load(n)[0]["world"]) == world]
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


def test_single_rank_and_grid_rules():
    sys.path.insert(0, ROOT)
    from arrow_matrix_b200.comm import SelfComm
    from arrow_matrix_b200.baseline import spmm_15d
    from tests.numpy_backend import GlooNumpyHaloFabric
    g, A = _load("p1_c1")
    lA, X, Y, grid, _, _, lNKb = spmm_15d.generate_15d_decomposition(A, int(g["k"]), np.float32, 1, None, comm=SelfComm(),
                                                                      X_full=g["X_full"])
    _check_rank(g, 0, lA, X, lNKb, grid)
    eng = spmm_15d.Spmm15D(grid, lA, X.shape[0], int(g["k"]), fabric=GlooNumpyHaloFabric(SelfComm()))
    eng.set_features(X)
    eng.spmm()
    assert np.allclose(eng.result(), g["r0_Y"], rtol=1e-5, atol=1e-6)
    # random features: generated on grid column 0 and replicated along the grid row
    _, X2, _, _, _, _, _ = spmm_15d.generate_15d_decomposition(A, 3, np.float32, 1, np.random.default_rng(0), comm=SelfComm())
    assert X2.shape == (A.shape[0], 3) and X2.dtype == np.float32

    class Fake(SelfComm):
        def __init__(self, size):
            self._n = size

        def Get_size(self):
            return self._n

    with pytest.raises(ValueError):
        spmm_15d.Grid15D(Fake(6), 4)                 # P not divisible by c (spmm_15d.py:34-36)
    with pytest.raises(ValueError):
        spmm_15d.Grid15D(Fake(8), 4)                 # P/c not divisible by c (:38-40)
    grid = spmm_15d.Grid15D(Fake(8), 2)
    assert (grid.p_div_c, grid.rounds) == (4, 2) and grid.Get_coords(5) == [2, 1] and grid.Get_cart_rank([2, 1]) == 5
    assert [spmm_15d.largest_power_of_two_square(p) for p in (1, 2, 4, 8, 16, 64)] == [1, 1, 2, 2, 4, 8]
    with pytest.raises(NotImplementedError):
        spmm_15d.spmm_15d_cpu(lA, X, Y, grid)


def test_driver_validates_and_logs(tmp_path, monkeypatch):
    """benchmark_15d on one rank with the numpy tiles: decomposition -> validation against A @ X -> timed products ->
    the reference's log artefacts (spmm_15d_main.py:155-281)"""
    sys.path.insert(0, ROOT)
    from arrow_matrix_b200 import synth, wb_logging
    from arrow_matrix_b200.comm import SelfComm
    from arrow_matrix_b200.baseline import spmm_15d
    from tests.numpy_backend import GlooNumpyHaloFabric
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(42)
    A = synth.generate_sparse_matrix(300, 300, 3000, np.float32, rng)
    out = spmm_15d.benchmark_15d(A, 8, 0, 3, "gpu", rng, comm=SelfComm(), verbose=False, fabric=GlooNumpyHaloFabric(SelfComm()))
    assert len(out["times"]) == 3 and out["validation"] < 1e-6
    runs = list(wb_logging.load_local_runs(tmp_path / "logs"))
    assert len(runs) == 1 and runs[0][0]["algorithm"] == "15D_B200_c_1_v0.1" and runs[0][0]["width"] == 300
    assert [e["iteration"] for e in runs[0][1] if "spmm_time" in e] == [0, 1, 2]
    trip = (A.data, A.indices, A.indptr)
    out = spmm_15d.benchmark_15d(trip, 4, 1, 1, "gpu", rng, new_decomposition=True, comm=SelfComm(), verbose=False,
                                 fabric=GlooNumpyHaloFabric(SelfComm()))
    assert out["validation"] < 1e-6
    with pytest.raises(NotImplementedError):
        spmm_15d.benchmark_15d(A, 4, 1, 1, "cpu", rng, comm=SelfComm(), verbose=False)
