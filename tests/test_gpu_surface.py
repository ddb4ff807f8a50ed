"""GPU tests written the way the reference's own tests read (tests/test_arrowmpi.py), through the public classes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle
from arrow_matrix_b200 import arrow_bench, graphio, synth
from arrow_matrix_b200.arrow_dec_mpi import ArrowDecompositionMPI
from arrow_matrix_b200.comm import SelfComm
from tests.test_gpu_kernels import assert_close


@pytest.mark.parametrize("b,k", [(2, 4), (5, 4), (9, 4), (64, 16), (100, 10)])
def test_decomposition_on_graph_flow(cuda_device, tmp_path, b, k):
    """mirrors test_decomposition_on_graph (tests/test_arrowmpi.py:205-309): save -> load -> initialize -> step ->
    _propagate_features -> allgather_result, golden = compute_spmm (tests/test_arrowdecomposition.py:139-156)"""
    comm = SelfComm()
    factor = 3
    dec = synth.synth_decomposition(factor, b, levels=2, perm_kind="random", seed=503, shrink=1)
    path = str(tmp_path / "test_ba")
    graphio.save_decomposition_new(dec, path, b, block_diagonal=True)
    blocks, n_blocks, to_prev, to_next = ArrowDecompositionMPI.load_decomposition_new(comm, path, b, True, slim=True)
    n = int(n_blocks[0]) * b
    arrow = ArrowDecompositionMPI.initialize(comm, n_blocks, to_prev, to_next, b, k, slim=True)
    permutations = [p for _, p in dec]
    rng = np.random.default_rng(42)
    X = np.round(rng.random((n, k), dtype=np.float32), 0)
    arrow.B.load_sparse_matrix_from_blocks(blocks)
    arrow.B.zero_rhs(b, k)
    assert arrow.matrix_index == 0 and arrow.B.is_column_rank() and arrow.decomposition_length == 2
    arrow.B.set_features_slice_from_features(X[permutations[0]])
    arrow.step()
    arrow._propagate_features()
    golden = oracle.compute_spmm(dec, X)
    C = np.zeros_like(golden, dtype=np.float32)
    arrow.B.allgather_result(C)
    assert np.allclose(C, golden[permutations[0]])
    # the deeper level holds golden_C in its own order after _propagate_features (test_arrowmpi.py:306-309)
    C1 = arrow.levels[1].C_i
    assert np.allclose(C1, golden[permutations[1]][: C1.shape[0]])


def test_single_arrow_matrix_spmm(cuda_device, tmp_path):
    """mirrors test_spmm (tests/test_arrowmpi.py:342-398): one arrow matrix, B.spmm(), allgather_result vs A @ X"""
    from scipy import sparse
    b, t, k = 2, 6, 1
    n = b * t
    rng = np.random.default_rng(42)
    A = np.zeros((n, n), dtype=np.float32)
    A[0:b, :] = rng.random((b, n), dtype=np.float32)
    A[:, 0:b] = rng.random((n, b), dtype=np.float32)
    A[1, 0:b] = 0
    A[:, 3] = 0
    for i in range(b, n):
        A[i, i] = rng.random()
    X = rng.random((n, k), dtype=np.float32)
    path = str(tmp_path / "one")
    graphio.save_decomposition_new([(sparse.csr_matrix(A), np.arange(n))], path, b, block_diagonal=True)
    comm = SelfComm()
    blocks, n_blocks, to_prev, to_next = ArrowDecompositionMPI.load_decomposition_new(comm, path, b, True, slim=True)
    arrow = ArrowDecompositionMPI.initialize(comm, n_blocks, to_prev, to_next, b, k, slim=True)
    arrow.B.load_sparse_matrix_from_blocks(blocks)
    arrow.B.zero_rhs(b, k)
    arrow.B.set_features(X)
    arrow.B.spmm()
    C = np.zeros_like(X)
    arrow.B.allgather_result(C)
    assert np.allclose(C, A @ X)
    assert np.allclose(arrow.B.feature_tile(), X)


def test_bench_spmm_driver_and_cli(cuda_device, tmp_path, monkeypatch):
    """mirrors test_larger_ranks (tests/test_arrowmpi.py:423-436) -- but checks the result, which the reference does not"""
    monkeypatch.chdir(tmp_path)
    out = arrow_bench.bench_spmm(None, 100, 2, 2, True, 'gpu', p_per_side=4, ba_neighbors=8, verbose=False)
    assert len(out["times"]) == 2
    out = arrow_bench.bench_spmm(None, 1000, 64, 3, True, 'gpu', p_per_side=4, ba_neighbors=9, slim=True, verbose=False)
    arrow = out["arrow"]
    dec = graphio.load_decomposition_new("tmp/test_ba_4_9", 1000, True)
    rng = np.random.default_rng(42)
    for _ in range(3):
        X = 2 * rng.random((4000, 64), dtype=np.float32) - 1        # the driver's last features (rng 42 + rank)
    got = arrow.B.result_tile()
    Xo = X[np.argsort(dec[0][1])]
    exact = oracle.compute_spmm([(B.astype(np.float64), p) for B, p in dec], Xo.astype(np.float64)).astype(np.float64)
    assert_close(oracle.to_original_order(got, dec[0][1], 4000), oracle.compute_spmm(dec, Xo), exact=exact)
    with pytest.raises(NotImplementedError):
        arrow_bench.bench_spmm(None, 100, 2, 1, True, 'cpu', p_per_side=2, verbose=False)
    from arrow_matrix_b200 import cli
    cli.main(["-w", "50", "-c", "8", "-z", "2", "-r", "3", "-m", "4"])


def test_bench_spmm_reference_route_ba_graph(cuda_device, tmp_path, monkeypatch):
    """the reference's own synthetic route (arrow_bench.py:33-34): Barabasi-Albert graph -> arrow decomposition ->
    files -> load -> iterate; the last level is best effort, so the check is against the reference protocol"""
    monkeypatch.chdir(tmp_path)
    out = arrow_bench.bench_spmm(None, 64, 8, 2, True, 'gpu', p_per_side=6, ba_neighbors=4, verbose=False, synthetic="ba")
    assert len(out["times"]) == 2
    arrow = out["arrow"]
    dec = graphio.load_decomposition_new("tmp/test_ba_6_4", 64, True)
    assert 1 <= len(dec) <= 3 and arrow.decomposition_length == len(dec)
    po = oracle.ReferenceProtocolOracle(dec, 64, 8)
    po64 = oracle.ReferenceProtocolOracle(dec, 64, 8, dtype=np.float64)
    rng = np.random.default_rng(42)
    for _ in range(2):                                   # the driver sets fresh features before every iteration
        X = 2 * rng.random((po.rows[0], 8), dtype=np.float32) - 1
        po.set_features(X.copy())
        po64.set_features(X)
        ref = po.step()
        exact = po64.step()
    assert_close(arrow.B.result_tile(), ref, exact=exact)
