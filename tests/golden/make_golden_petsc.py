"""Golden vectors of the PETSc-style 1D baseline, produced by the UNMODIFIED reference (no GPU, no real MPI).

    python tests/golden/make_golden_petsc.py      # rewrites tests/golden/petsc_*.npz

Runs the reference's ``MatrixSlice.initialize`` (``arrow/matrix_slice.py:107-156``) and ``spmm_cpu``
(``arrow/baseline/spmm_petsc.py:183-226``) with one thread per MPI rank (``fake_mpi.py``), the way the reference's
own ``tests/test_spmmPETSc.py`` drives them, and stores the inputs, every rank's communication tables and result.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import fake_mpi  # noqa: E402

fake_mpi.install()
sys.path.insert(0, "/root/reference")

from arrow.matrix_slice import MatrixSlice  # noqa: E402  (the reference)
from arrow.baseline.spmm_petsc import spmm_cpu  # noqa: E402  (the reference)


def _rank_main(comm, A, X, bounds):
    r = comm.Get_rank()
    s, e = int(bounds[r]), int(bounds[r + 1])
    A_mine = A[s:e, :]
    X_i = np.array(X[s:e, :])
    sl = MatrixSlice.initialize(comm, A_mine)
    Y_i = np.zeros((e - s, X.shape[1]), dtype=X.dtype)
    X_non = np.zeros((len(sl.rank_in), X.shape[1]), dtype=X.dtype)
    comm.Barrier()
    Y = spmm_cpu(comm, sl, X_i, Y_i, X_non)
    loc, non = scipy.sparse.csr_matrix(sl.A_i_local), scipy.sparse.csr_matrix(sl.A_i_nonlocal)
    loc.sort_indices()
    non.sort_indices()
    return dict(x_index_in=np.asarray(sl.x_index_in), rank_in=np.asarray(sl.rank_in), x_index_out=np.asarray(sl.x_index_out),
                rank_out=np.asarray(sl.rank_out), send_count=np.asarray(sl.send_count), recv_count=np.asarray(sl.recv_count),
                x_index_out_localized=np.asarray(sl.x_index_out_localized),
                x_index_in_localized=np.asarray(sl.x_index_in_localized),
                send_sdispl=np.asarray(sl.send_sdispl), recv_sdispl=np.asarray(sl.recv_sdispl),
                start_col=np.int64(sl.start_col), end_col=np.int64(sl.end_col),
                loc_shape=np.asarray(loc.shape), loc_indptr=loc.indptr, loc_indices=loc.indices, loc_data=loc.data,
                non_shape=np.asarray(non.shape), non_indptr=non.indptr, non_indices=non.indices, non_data=non.data,
                X_nonlocal=X_non, Y=np.asarray(Y))


def save_case(name, A, X, all_n_i, note):
    A = scipy.sparse.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    A.eliminate_zeros()                                  # the reference's driver does (spmm_petsc.py:431-433)
    world = len(all_n_i)
    bounds = np.concatenate([[0], np.cumsum(all_n_i)]).astype(np.int64)
    res = fake_mpi.run_world(world, _rank_main, A, X, bounds)
    out = dict(note=note, world=world, all_n_i=np.asarray(all_n_i, dtype=np.int64), A_indptr=A.indptr, A_indices=A.indices,
               A_data=A.data, n=A.shape[0], X=X)
    for r, d in enumerate(res):
        for key, v in d.items():
            out[f"r{r}_{key}"] = v
    full = A @ X
    for r in range(world):                               # the reference test's own assertion (test_spmmPETSc.py:36-43)
        assert np.allclose(res[r]["Y"], full[bounds[r]:bounds[r + 1]]), (name, r)
    np.savez_compressed(os.path.join(HERE, f"petsc_{name}.npz"), **out)
    print(name, "world", world, "n", A.shape[0], "halo rows", [int(d["x_index_in"].size) for d in res])


def main():
    rng = np.random.default_rng(11)
    k = 4
    # unequal slices incl. an empty one (test_spmm_unequal, test_spmmPETSc.py:45-72)
    for world, sizes in ((2, [33, 7]), (3, [33, 33, 0]), (4, [20, 0, 13, 31])):
        n = int(sum(sizes))
        A = scipy.sparse.rand(n, n, density=0.05, format="csr", random_state=42 + world, dtype=np.float64).astype(np.float32)
        X = np.round(5 * rng.random((n, k))).astype(np.float32)
        save_case(f"unequal_w{world}", A, X, sizes, "random density 0.05, integer-valued X (exact sums)")
    # identity: nothing crosses ranks (test_spmm_eye, :93-118)
    A = scipy.sparse.eye(48, dtype=np.float32, format="csr")
    save_case("eye_w3", A, rng.random((48, 8)).astype(np.float32), [16, 16, 16], "identity, no halo")
    # empty matrix (density 0 in test_spmm_unequal)
    A = scipy.sparse.csr_matrix((24, 24), dtype=np.float32)
    save_case("empty_w2", A, rng.random((24, 4)).astype(np.float32), [12, 12], "no stored entry")
    # denser, float values, one rank
    A = scipy.sparse.rand(40, 40, density=0.2, format="csr", random_state=5, dtype=np.float64).astype(np.float32)
    save_case("single_w1", A, (2 * rng.random((40, 6)) - 1).astype(np.float32), [40], "world of one")
    # power-law-ish columns (hub columns requested by every rank)
    n = 120
    rows = rng.integers(0, n, size=900)
    cols = np.minimum((rng.pareto(1.2, size=900) * 3).astype(np.int64), n - 1)
    A = scipy.sparse.csr_matrix((rng.random(900).astype(np.float32), (rows, cols)), shape=(n, n))
    save_case("hubs_w4", A, (2 * rng.random((n, 16)) - 1).astype(np.float32), [30, 30, 30, 30], "hub columns")


if __name__ == "__main__":
    main()
