"""Generate golden vectors by running the UNMODIFIED reference here (no GPU, no real MPI).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference classes are imported from /root/reference with ``mpi4py`` replaced by the
thread-based stand-in in ``fake_mpi.py`` (one thread per MPI rank, one rank per block-row exactly
as ``arrow/arrow_bench.py:70-78`` demands).  Everything between the npy files and ``B.C_i`` is the
reference's own code: ``ArrowDecompositionMPI.load_decomposition_new`` (root read + scatter),
``initialize`` (communicators + all-to-all tables), ``ArrowSlimMPI`` / ``ArrowMPI`` ``spmm`` (SciPy
arithmetic) and ``step()``.  The fixtures hold the inputs (CSR triplets, permutations, features) and the
per-level result tiles after every iteration, so the tests need neither /root/reference nor MPI.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import fake_mpi  # noqa: E402

fake_mpi.install()
sys.path.insert(0, "/root/reference")

from arrow.arrow_dec_mpi import ArrowDecompositionMPI  # noqa: E402  (the reference)

from arrow_matrix_b200 import graphio, synth  # noqa: E402  (only to write input files)


def _rank_main(comm, path, width, k, slim, block_diagonal, Xs):
    blocks, n_blocks, to_prev, to_next = ArrowDecompositionMPI.load_decomposition_new(
        comm, path, width, block_diagonal, np.float32, slim=slim, use_npy=True, use_mmap=False)
    arrow = ArrowDecompositionMPI.initialize(comm, n_blocks, to_prev, to_next, width, k, 'cpu', block_diagonal, slim)
    if arrow is None:
        return None
    arrow.B.load_sparse_matrix_from_blocks(blocks)
    arrow.B.zero_rhs(width, k)
    is_col = arrow.B.is_column_rank()
    col_rank = arrow.B.column_comm.Get_rank() if is_col else -1
    tiles = []
    for X in Xs:
        if X is not None and arrow.matrix_index == 0 and is_col:
            arrow.B.set_features(np.array(X[col_rank * width:(col_rank + 1) * width], dtype=np.float32))
        arrow.comm.Barrier()
        arrow.step()
        tiles.append(np.array(arrow.B.C_i, dtype=np.float32) if is_col else None)
    arrow._propagate_features()
    final = np.array(arrow.B.C_i, dtype=np.float32) if is_col else None
    return dict(level=arrow.matrix_index, col_rank=col_rank, tiles=tiles, final=final,
                n_blocks=np.array(n_blocks), to_prev=to_prev, to_next=to_next)


def run_reference(dec, width, k, slim, block_diagonal, Xs, write_kwargs=None):
    """-> n_blocks, C[it][level] (concatenated column-rank tiles), final[level], to_prev[level], to_next[level]."""
    from oracle import oracle as _o
    nb = [_o.number_of_blocks(B, width) for B, _ in dec]
    n_ranks = sum(nb) if slim else sum(2 * b - 1 for b in nb)
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "g")
        graphio.save_decomposition_new(dec, base, width, block_diagonal, **(write_kwargs or {}))
        res = fake_mpi.run_world(n_ranks, _rank_main, base, width, k, slim, block_diagonal, Xs)
    L = len(dec)
    n_blocks = res[0]["n_blocks"]
    assert list(n_blocks) == nb
    C = [[np.zeros((nb[j] * width, k), np.float32) for j in range(L)] for _ in Xs]
    final = [np.zeros((nb[j] * width, k), np.float32) for j in range(L)]
    to_prev = [np.full(nb[j] * width, -1, np.int64) for j in range(L)]
    to_next = [np.full(nb[j] * width, -1, np.int64) for j in range(L)]
    for r in res:
        if r is None or r["col_rank"] < 0:
            continue
        j, c = r["level"], r["col_rank"]
        sl = slice(c * width, (c + 1) * width)
        for it in range(len(Xs)):
            C[it][j][sl] = r["tiles"][it]
        final[j][sl] = r["final"]
        if r["to_prev"] is not None:
            to_prev[j][sl] = r["to_prev"]
        if r["to_next"] is not None:
            to_next[j][sl] = r["to_next"]
    return n_blocks, C, final, to_prev, to_next


def save_case(name, dec, width, k, slim, block_diagonal, Xs, write_kwargs=None, note=""):
    n_blocks, C, final, to_prev, to_next = run_reference(dec, width, k, slim, block_diagonal, Xs, write_kwargs)
    out = dict(width=width, k=k, slim=slim, block_diagonal=block_diagonal, levels=len(dec), iterations=len(Xs),
               n_blocks=np.asarray(n_blocks), note=note,
               write_data=(write_kwargs or {}).get("write_data", True),
               one_based=(write_kwargs or {}).get("one_based_permutation", False))
    for j, (B, p) in enumerate(dec):
        out[f"indptr_{j}"], out[f"indices_{j}"], out[f"data_{j}"] = B.indptr, B.indices, B.data
        out[f"perm_{j}"] = np.asarray(p)
        out[f"final_{j}"] = final[j]
        out[f"to_prev_{j}"], out[f"to_next_{j}"] = to_prev[j], to_next[j]
        for it in range(len(Xs)):
            out[f"C_{it}_{j}"] = C[it][j]
    for it, X in enumerate(Xs):
        out[f"X_{it}"] = np.zeros((0, k), np.float32) if X is None else X
        out[f"has_X_{it}"] = X is not None
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: n_blocks={list(n_blocks)} ranks={'slim' if slim else 'wide'} k={k}")


def tables_kats():
    """Known answers of the reference's static ``_all_to_all_tables`` on seeded inputs."""
    rng = np.random.default_rng(2024)
    cases = {}
    idx = 0
    for rpr, total, cols in [(4, 8, 6), (8, 3, 3), (16, 5, 2), (10, 7, 1)]:
        for off in (0, total // 2):
            perm = rng.permutation(rpr * total)[:rpr].astype(np.int64)
            perm[rng.integers(0, rpr)] = 2 * rpr * total           # a sentinel
            c, d, sp, rp = ArrowDecompositionMPI._all_to_all_tables(perm, rpr, cols, total, off)
            cases[f"in_{idx}"] = np.concatenate([[rpr, cols, total, off], perm])
            cases[f"counts_{idx}"], cases[f"displs_{idx}"] = np.asarray(c, np.int64), np.asarray(d, np.int64)
            cases[f"send_{idx}"], cases[f"recv_{idx}"] = np.asarray(sp, np.int64), np.asarray(rp, np.int64)
            idx += 1
    cases["n"] = idx
    np.savez_compressed(os.path.join(HERE, "all_to_all_tables.npz"), **cases)
    print(f"wrote all_to_all_tables.npz ({idx} cases)")


def main():
    rng = np.random.default_rng(7)

    def feats(n, k, integer=False):
        X = synth.generate_dense_matrix(n, k, np.float32, rng)
        return np.round(4 * X).astype(np.float32) if integer else X

    tables_kats()

    # A: slim, 2 levels, random permutation; iteration 2 is chained (no new features), 3 gets fresh ones
    w, t0, k = 8, 5, 4
    dec = synth.synth_decomposition(t0, w, levels=2, perm_kind="random", seed=101)
    save_case("slim_L2_random_k4", dec, w, k, True, True, [feats(t0 * w, k), None, feats(t0 * w, k)],
              note="iteration 1 is chained: X := A X from iteration 0")

    # B: slim, 3 levels, NON-nested permutations -> sentinel / stale rows; last level is a single rank
    w, t0, k = 6, 6, 3
    dec = synth.synth_decomposition(t0, w, levels=3, perm_kind="random", seed=202, nested=False)
    save_case("slim_L3_nonnested_k3", dec, w, k, True, True, [feats(t0 * w, k), feats(t0 * w, k), feats(t0 * w, k)],
              note="rows behind the sentinel keep stale values (arrow_dec_mpi.py:544)")

    # C: slim, hub rows in the head, k=16, shard-local permutation, integer-valued features (exact sums)
    w, t0, k = 16, 4, 16
    dec = synth.synth_decomposition(t0, w, levels=2, perm_kind="local", seed=303, hub_rows=3, hub_nnz=40)
    save_case("slim_L2_hubs_k16", dec, w, k, True, True, [feats(t0 * w, k, True), feats(t0 * w, k, True)])

    # D: wide layout (ArrowMPI, 2t-1 ranks per level), block diagonal
    w, t0, k = 8, 4, 5
    dec = synth.synth_decomposition(t0, w, levels=2, perm_kind="random", seed=404)
    save_case("wide_L2_random_k5", dec, w, k, False, True, [feats(t0 * w, k), None])

    # G: wide layout, BANDED (non block-diagonal): blocks (i, i-1), (i, i+1) and the neighbour tile exchange
    #    (arrow_mpi.py:123-175, 211-219)
    w, t0, k = 8, 5, 4
    dec = synth.synth_decomposition(t0, w, levels=2, perm_kind="random", seed=707, band_nnz=3, shrink=1)
    save_case("wide_L2_banded_k4", dec, w, k, False, False, [feats(t0 * w, k), feats(t0 * w, k)],
              note="non block-diagonal: A_i,i-1 and A_i,i+1 present")

    # E: Julia-converter quirks: no data file (ones), int64 indices, 1-based permutations
    w, t0, k = 8, 4, 4
    dec = synth.synth_decomposition(t0, w, levels=2, perm_kind="random", seed=505)
    dec1 = [(B.__class__((np.ones_like(B.data), B.indices, B.indptr), shape=B.shape), p) for B, p in dec]
    save_case("slim_L2_julia_quirks_k4", dec1, w, k, True, True, [feats(t0 * w, k)],
              write_kwargs=dict(write_data=False, index_dtype=np.int64, one_based_permutation=True))

    # F: a level-0 file shorter than n_blocks*width (padding of permutations and of the last block-row)
    w, t0, k = 8, 4, 4
    dec = synth.synth_decomposition(t0, w, levels=2, perm_kind="random", seed=606)
    n_short = t0 * w - 3
    dec2 = []
    for B, p in dec:
        Bs = B[:n_short, :n_short].tocsr()
        ps = np.asarray([x for x in p if x < n_short])
        dec2.append((Bs, ps))
    save_case("slim_L2_short_file_k4", dec2, w, k, True, True, [feats(t0 * w, k)],
              note="features cover n_blocks[0]*width rows; rows beyond the file are padding")

    # H: slim, FOUR levels, k not a multiple of 4, three chained iterations (new cases go last: the feature stream of
    #    the earlier cases must not move)
    w, t0, k = 4, 8, 6
    dec = synth.synth_decomposition(t0, w, levels=4, perm_kind="random", seed=808)
    save_case("slim_L4_nested_k6", dec, w, k, True, True, [feats(t0 * w, k), None, None],
              note="four levels; iterations 1 and 2 are chained")

    # I: wide layout, banded, THREE levels with non-nested permutations (stale rows + neighbour tiles together)
    w, t0, k = 6, 5, 7
    dec = synth.synth_decomposition(t0, w, levels=3, perm_kind="random", seed=909, nested=False, band_nnz=2, shrink=1)
    save_case("wide_L3_banded_stale_k7", dec, w, k, False, False, [feats(t0 * w, k), feats(t0 * w, k), None],
              note="non block-diagonal, three levels, rows behind the sentinel keep stale values")


if __name__ == "__main__":
    main()
