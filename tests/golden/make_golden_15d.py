"""Golden vectors of the A-stationary 1.5D baseline, produced by the UNMODIFIED reference (no GPU, no real MPI).

    python tests/golden/make_golden_15d.py      # rewrites tests/golden/spmm15d_*.npz

Runs the reference's ``generate_15d_decomposition`` (``arrow/baseline/spmm_15d.py:19-154``) and ``spmm_15d_cpu``
(``:313-367``) with one thread per MPI rank on a (P/c) x c grid (``fake_mpi.py``) and stores, per rank, its A blocks,
its X block and its result.
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

from arrow.baseline.spmm_15d import generate_15d_decomposition, spmm_15d_cpu  # noqa: E402  (the reference)


def _rank_main(comm, A, k, c, seed):
    rng = np.random.default_rng(seed + comm.Get_rank())
    lA, X, Y, cart, bc, rd, lNKb = generate_15d_decomposition(A if comm.Get_rank() == 0 else None, k, np.float32, c, rng)
    Y = spmm_15d_cpu(lA, X, Y, cart, bc, rd)
    out = dict(X=np.array(X), Y=np.array(Y), lNKb=np.int64(lNKb), coords=np.asarray(cart.Get_coords(cart.Get_rank())),
               n_blocks=np.int64(len(lA)))
    for r, B in enumerate(lA):
        B = scipy.sparse.csr_matrix(B)
        out[f"A{r}_shape"] = np.asarray(B.shape)
        out[f"A{r}_indptr"], out[f"A{r}_indices"], out[f"A{r}_data"] = B.indptr, B.indices, B.data
    return out


def save_case(name, A, k, world, c, seed, note):
    A = scipy.sparse.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    res = fake_mpi.run_world(world, _rank_main, A, k, c, seed)
    out = dict(note=note, world=world, c=c, k=k, n=A.shape[0], A_indptr=A.indptr, A_indices=A.indices, A_data=A.data)
    for r, d in enumerate(res):
        for key, v in d.items():
            out[f"r{r}_{key}"] = v
    # reassemble X from the ranks of grid column 0 and check Y = A X on every rank (replicated along the grid row)
    p_div_c = world // c
    lNKb = int(res[0]["lNKb"])
    X = np.concatenate([res[x * c]["X"] for x in range(p_div_c)])[: A.shape[1]]
    full = A @ X
    lNI = int(np.ceil(A.shape[0] / p_div_c))
    for r, d in enumerate(res):
        x = r // c
        assert np.allclose(d["Y"], full[x * lNI:(x + 1) * lNI], rtol=1e-4, atol=1e-5), (name, r)
    out["X_full"] = X
    np.savez_compressed(os.path.join(HERE, f"spmm15d_{name}.npz"), **out)
    print(name, "world", world, "c", c, "rounds", p_div_c // c, "lNKb", lNKb, "Y rows", [int(d["Y"].shape[0]) for d in res])


def main():
    def rand(n, density, seed):
        return scipy.sparse.rand(n, n, density=density, format="csr", random_state=seed, dtype=np.float64).astype(np.float32)
    save_case("p1_c1", rand(37, 0.1, 1), 4, 1, 1, 5, "one rank")
    save_case("p2_c1", rand(50, 0.08, 2), 6, 2, 1, 6, "1D: two block rows, two rounds")
    save_case("p4_c1", rand(61, 0.06, 3), 4, 4, 1, 7, "1D: four rounds, ragged last block")
    save_case("p4_c2", rand(64, 0.06, 4), 8, 4, 2, 8, "2 x 2 grid, one round, all-reduce over 2")
    save_case("p8_c2", rand(90, 0.05, 5), 5, 8, 2, 9, "4 x 2 grid, two rounds")
    save_case("p4_c2_ragged", rand(35, 0.1, 6), 4, 4, 2, 10, "2 x 2 grid, n not divisible")


if __name__ == "__main__":
    main()
