"""CPU tests of the igraph-free arrow decomposition, with the properties the reference's own test asserts
(tests/test_arrowdecomposition.py:24-112): edges are partitioned, permutations are permutations, non-final levels
respect the width, sum_j P_j B_j P_j^T == A, and the decomposed product equals A @ X."""
import numpy as np
import pytest
from scipy import sparse

from arrow_matrix_b200 import graphio, synth
from arrow_matrix_b200.decomposition import arrow_decomposition, reconstruct
from oracle import oracle


def erdos_renyi(n, p, seed):
    rng = np.random.default_rng(seed)
    U = sparse.random(n, n, density=p / 2, format="coo", random_state=rng, dtype=np.float32)
    A = sparse.csr_matrix(((U + U.T) > 0).astype(np.float32))
    A.setdiag(0)
    A.eliminate_zeros()
    return A


GRAPHS = [("ba16", lambda: synth.barabasi_albert(16, 3)), ("ba512", lambda: synth.barabasi_albert(512, 5)),
          ("ba4096", lambda: synth.barabasi_albert(4096, 3)), ("er300", lambda: erdos_renyi(300, 0.02, 1)),
          ("er_sparse_with_isolated", lambda: erdos_renyi(400, 0.004, 2))]


@pytest.mark.parametrize("block_diagonal", [True, False])
@pytest.mark.parametrize("name,make", GRAPHS)
@pytest.mark.parametrize("width_count", [2, 5])
def test_arrow_properties(name, make, width_count, block_diagonal):
    A = make()
    n = A.shape[0]
    width = n // width_count + 1
    dec = arrow_decomposition(A, width, max_number_of_levels=4, block_diagonal=block_diagonal, seed=42)
    assert 1 <= len(dec) <= 4
    total_nnz = 0
    for j, (B, perm) in enumerate(dec):
        assert sorted(perm.tolist()) == list(range(n))                      # a permutation
        assert B.shape == (n, n) and B.has_canonical_format
        total_nnz += B.nnz
        C = B.tocoo()
        if j < len(dec) - 1:                                                # width criterion of non-final levels
            if block_diagonal:
                ok = (C.row // width == C.col // width) | (C.row < width) | (C.col < width)
            else:
                ok = (np.abs(C.row - C.col) <= width) | (C.row < width) | (C.col < width)
            assert np.all(ok)
        assert abs(B - B.T).nnz == 0                                        # levels stay symmetric
    assert total_nnz == A.nnz                                               # the levels partition the edges
    assert abs(reconstruct(dec, n) - A.astype(np.float64)).max() < 1e-6      # sum_j P_j B_j P_j^T == A
    rng = np.random.default_rng(42)
    X = rng.random((n, 16), dtype=np.float32)
    assert np.allclose(oracle.compute_spmm(dec, X), A @ X, rtol=1e-4, atol=1e-4)


def test_decomposition_feeds_the_engine_format(tmp_path):
    """decompose -> save in the npy layout -> load -> the chained protocol reproduces A @ X (when nothing is dropped)"""
    A = synth.barabasi_albert(600, 4, seed=7)
    n, w, k = 600, 100, 8
    dec = arrow_decomposition(A, w, max_number_of_levels=3, block_diagonal=True, seed=1)
    base = str(tmp_path / "g")
    graphio.save_decomposition_new(dec, base, w, block_diagonal=True)
    back = graphio.load_decomposition_new(base, w, True)
    assert len(back) == len(dec)
    po = oracle.ReferenceProtocolOracle(back, w, k)
    X = np.random.default_rng(0).random((n, k), dtype=np.float32)
    Xp = np.zeros((po.rows[0], k), np.float32)
    Xp[:n] = X[po.perms[0][:n]] if po.rows[0] >= n else X[po.perms[0]]
    po.set_features(Xp)
    C0 = po.step()
    got = oracle.to_original_order(C0, po.perms[0], n)
    ref = A @ X
    if sum(po.dropped_nnz) == 0 and all(np.all(po.to_prev[j][: po.rows[j]] < po.rows[j - 1]) for j in range(1, po.L)):
        assert np.allclose(got, ref, rtol=1e-4, atol=1e-4)
    else:       # the reference drops what falls outside the arrow pattern of the best-effort level (graphio.py:382-383)
        assert got.shape == ref.shape


def test_head_holds_the_hubs():
    A = synth.barabasi_albert(2000, 3, seed=3)
    dec = arrow_decomposition(A, 50, max_number_of_levels=2, block_diagonal=True, seed=0)
    deg = np.diff(A.indptr)
    head = dec[0][1][:50]
    assert set(head.tolist()) == set(np.argsort(-deg, kind="stable")[:50].tolist())


def test_arrow_decompose_cli_roundtrip(tmp_path):
    """mtx file -> arrow_decompose -> npy layout -> load -> sum_j P_j B_j P_j^T == A"""
    from scipy import io as sio
    from arrow_matrix_b200 import decompose_cli
    A = synth.barabasi_albert(300, 3, seed=11)
    d = tmp_path / "toy"
    d.mkdir()
    sio.mmwrite(str(d / "toy.mtx"), sparse.triu(A))           # upper triangle only: the CLI symmetrises
    decompose_cli.main(["--width", "40", "--dataset_dir", str(tmp_path), "--dataset_name", "toy", "--format", "mtx"])
    dec = graphio.load_decomposition_new(str(d / "toy"), 40, True)
    assert len(dec) >= 1
    assert abs(reconstruct(dec, 300) - A.astype(np.float64)).max() < 1e-6
