"""CPU tests: oracle and host-side loader logic against golden vectors from the REAL reference.

The fixtures in tests/golden/*.npz were produced by tests/golden/make_golden.py, which runs the
unmodified /root/reference classes (load_decomposition_new -> initialize -> step) under an in-process
MPI stand-in.  This is what pins the oracle ("parity pinned").
"""
import numpy as np
import pytest

from oracle import oracle
from arrow_matrix_b200 import decomp
from tests.golden_util import CASES, GOLDEN_DIR, GoldenCase


def test_fixtures_present():
    assert len(CASES) >= 6


def test_all_to_all_tables_against_reference_outputs():
    import os
    z = np.load(os.path.join(GOLDEN_DIR, "all_to_all_tables.npz"))
    for i in range(int(z["n"])):
        head = z[f"in_{i}"]
        rpr, cols, total, off = (int(x) for x in head[:4])
        perm = head[4:]
        c, d, sp, rp = oracle.all_to_all_tables(perm, rpr, cols, total, off)
        assert np.array_equal(c, z[f"counts_{i}"]) and np.array_equal(d, z[f"displs_{i}"])
        assert np.array_equal(sp, z[f"send_{i}"]) and np.array_equal(rp, z[f"recv_{i}"])


@pytest.mark.parametrize("name", CASES)
def test_loader_semantics_match_reference(name):
    g = GoldenCase(name)
    nb_o = [oracle.number_of_blocks(B, g.width) for B, _ in g.decomposition]
    nb_p = [decomp.number_of_blocks(B, g.width) for B, _ in g.decomposition]
    assert nb_o == g.n_blocks and nb_p == g.n_blocks
    for mod in (oracle, decomp):
        perms, to_prev, to_next, sentinel = mod.prepare_permutations([p for _, p in g.decomposition], g.n_blocks, g.width)
        assert sentinel == 2 * g.width * g.n_blocks[0]
        for j in range(g.L):
            rows = g.n_blocks[j] * g.width
            if j > 0:
                assert np.array_equal(to_prev[j][:rows], g.to_prev[j]), (name, j)
            if j < g.L - 1:
                assert np.array_equal(to_next[j][:rows], g.to_next[j]), (name, j)


@pytest.mark.parametrize("blockwise", [False, True])
@pytest.mark.parametrize("use_c", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_protocol_oracle_reproduces_reference_run(name, use_c, blockwise):
    g = GoldenCase(name)
    po = oracle.ReferenceProtocolOracle(g.decomposition, g.width, g.k, block_diagonal=g.block_diagonal,
                                        use_c_kernel=use_c, blockwise=blockwise)
    assert po.n_blocks == g.n_blocks
    for it in range(g.iterations):
        if g.X[it] is not None:
            po.set_features(g.X[it].copy())
        po.step()
        for j in range(g.L):
            assert np.allclose(po.C[j], g.C[it][j], rtol=1e-5, atol=1e-6), (name, it, j)
    po.propagate_features()
    for j in range(g.L):
        assert np.allclose(po.C[j], g.final[j], rtol=1e-5, atol=1e-6), (name, "final", j)


@pytest.mark.parametrize("name", CASES)
def test_arrow_rows_equals_oracle_mask(name):
    """product-side block restriction (decomp.arrow_rows) == oracle.arrow_mask, full range and row shards"""
    from scipy import sparse
    g = GoldenCase(name)
    for (B, _), nb in zip(g.decomposition, g.n_blocks):
        M = oracle.arrow_mask(B, g.width, nb, g.block_diagonal)
        n = nb * g.width
        for (r0, r1) in [(0, n), (g.width, n), (0, g.width), (n // 2, n)]:
            ip, idx, dat, dropped = decomp.arrow_rows(B, g.width, nb, g.block_diagonal, r0, r1, chunk_rows=7)
            S = sparse.csr_matrix((dat, idx, ip), shape=(r1 - r0, n))
            assert (abs(S - M[r0:r1]) > 0).nnz == 0
