"""Row slice of a 1D-partitioned sparse matrix plus the halo-exchange tables of its SpMM (SURVEY.md "next" row N4).

Same class name, attributes and table conventions as the reference's ``arrow/matrix_slice.py:10-290`` -- this is the
set-up half of the PETSc-style comparison baseline (``arrow/baseline/spmm_petsc.py``):

* rank ``i`` holds ``A_i``: ``n_i`` consecutive rows of a square matrix, all columns; the ``n_i`` are all-gathered and
  the rank's *local* columns are ``[start_col, end_col) = [sum n_{<i}, sum n_{<=i})`` (``:83-104``);
* ``A_i_local = A_i[:, start_col:end_col]``; ``A_i_nonlocal`` = the non-empty columns outside that range, compacted
  in ascending column order (``:125-147``);
* ``x_index_in`` = those global column ids (sorted), ``rank_in`` = the rank owning each (``:185-227``);
* the owners learn what to send through an all-to-all of the counts and of the index lists:
  ``x_index_out`` / ``rank_out`` sorted by ``(rank_out, x_index_out)`` (``:234-290``);
* ``send_count/recv_count`` per peer, ``send_sdispl/recv_sdispl`` their exclusive prefix sums (one entry longer),
  ``x_index_out_localized = x_index_out - start_col`` and ``x_index_in_localized`` = the row inside the owner's slice.

The reference builds the tables with Python loops over columns (``:207-219``, ``:270-282``); here they are vectorised
(``searchsorted`` / stable sorts), which matters at millions of halo rows.  The communicator only needs ``allgather``
and ``alltoall`` of Python objects (``arrow_matrix_b200.comm``).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy import sparse


class MatrixSlice:
    def __init__(self, A_i_local: sparse.csr_matrix, A_i_nonlocal: sparse.csr_matrix, x_index_in: np.ndarray,
                 rank_in: np.ndarray, x_index_out: np.ndarray, rank_out: np.ndarray, all_n_i: np.ndarray,
                 start_col: int, end_col: int, send_count: np.ndarray, recv_count: np.ndarray):
        assert A_i_local.shape[0] == A_i_nonlocal.shape[0]
        assert x_index_in.shape == rank_in.shape and x_index_out.shape == rank_out.shape
        assert A_i_nonlocal.shape[1] == x_index_in.size
        self.A_i_local = A_i_local
        self.A_i_nonlocal = A_i_nonlocal
        self.x_index_in = x_index_in
        self.rank_in = rank_in
        self.x_index_out = x_index_out
        self.rank_out = rank_out
        self.all_n_i = np.asarray(all_n_i, dtype=np.int64)
        self.start_col = int(start_col)
        self.end_col = int(end_col)
        self.x_index_out_localized = x_index_out - self.start_col
        assert np.all(self.x_index_out_localized >= 0) and np.all(self.x_index_out_localized < self.end_col - self.start_col)
        starts = np.concatenate([[0], np.cumsum(self.all_n_i)[:-1]]).astype(np.int64)
        self.x_index_in_localized = x_index_in - starts[rank_in] if rank_in.size else x_index_in.copy()
        self.send_count = np.asarray(send_count, dtype=np.int64)
        self.recv_count = np.asarray(recv_count, dtype=np.int64)
        self.send_sdispl = np.concatenate([[0], np.cumsum(self.send_count)]).astype(np.int64)
        self.recv_sdispl = np.concatenate([[0], np.cumsum(self.recv_count)]).astype(np.int64)

    # -- pieces of initialize(), kept as separate static methods like the reference -------------------------
    @staticmethod
    def get_local_matrix_dimensions(comm, A_i) -> np.ndarray:
        return np.asarray(comm.allgather(int(A_i.shape[0])), dtype=np.int64)

    @staticmethod
    def identify_local_slice(rank: int, all_n_i) -> Tuple[int, int]:
        all_n_i = np.asarray(all_n_i, dtype=np.int64)
        return int(all_n_i[:rank].sum()), int(all_n_i[:rank + 1].sum())

    @staticmethod
    def _is_sorted(a) -> bool:
        a = np.asarray(a)
        return bool(np.all(a[:-1] <= a[1:]))

    @staticmethod
    def construct_receive_tables(A_i, start_col: int, end_col: int, all_n_i) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """``(nonlocal_cols, x_index_in, rank_in)``, sorted by column: the columns with a stored entry outside the
        rank's own range and their owners (``:185-227``; stored zeros count, like ``nonzero()`` after
        ``eliminate_zeros`` in the reference's driver, ``spmm_petsc.py:431``)."""
        A_i = sparse.csr_matrix(A_i)
        cols = A_i.indices[A_i.data != 0] if A_i.nnz else A_i.indices
        outside = cols[(cols < start_col) | (cols >= end_col)]
        nonlocal_cols = np.unique(outside).astype(np.int64)
        bounds = np.cumsum(np.asarray(all_n_i, dtype=np.int64))
        rank_in = np.searchsorted(bounds, nonlocal_cols, side="right").astype(np.int64)
        return nonlocal_cols, nonlocal_cols.copy(), rank_in

    @staticmethod
    def construct_send_tables(comm, rank_in: np.ndarray, x_index_in: np.ndarray, recv_counts: np.ndarray
                              ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """``(x_index_out, rank_out, send_counts)``: what every peer asked this rank for (``:234-290``)."""
        assert MatrixSlice._is_sorted(x_index_in) and MatrixSlice._is_sorted(rank_in)
        size = comm.Get_size()
        bounds = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
        wanted = comm.alltoall([np.asarray(x_index_in[bounds[g]:bounds[g + 1]], dtype=np.int64) for g in range(size)])
        send_counts = np.array([w.size for w in wanted], dtype=np.int64)
        x_index_out = np.concatenate(wanted) if size else np.zeros(0, np.int64)        # each list arrives sorted
        rank_out = np.repeat(np.arange(size, dtype=np.int64), send_counts)
        return x_index_out.astype(np.int64), rank_out, send_counts

    @staticmethod
    def check_comm_tables(comm, x_index_in, rank_in, x_index_out, rank_out) -> bool:
        """Every (index, requester) pair a rank will serve was requested by that peer, and nothing else."""
        me = comm.Get_rank()
        asked = comm.allgather((np.asarray(x_index_in), np.asarray(rank_in)))
        for g, (idx, owner) in enumerate(asked):
            want = idx[owner == me]
            if not np.array_equal(np.sort(want), np.sort(np.asarray(x_index_out)[np.asarray(rank_out) == g])):
                return False
        return True

    @classmethod
    def initialize(cls, comm, A_i) -> "MatrixSlice":
        rank = comm.Get_rank()
        A_i = sparse.csr_matrix(A_i)
        all_n_i = cls.get_local_matrix_dimensions(comm, A_i)
        total_rows = int(all_n_i.sum())
        if total_rows != A_i.shape[1]:
            raise ValueError(f"Matrix not square: Rank {rank} has {A_i.shape[1]} columns, "
                             f"but the total number of rows is {total_rows}")
        start_col, end_col = cls.identify_local_slice(rank, all_n_i)
        A_i_local = sparse.csr_matrix(A_i[:, start_col:end_col])
        A_i_local.sum_duplicates()
        A_i_local.sort_indices()
        A_i_local.eliminate_zeros()
        nonlocal_cols, x_index_in, rank_in = cls.construct_receive_tables(A_i, start_col, end_col, all_n_i)
        recv_counts = np.bincount(rank_in, minlength=comm.Get_size()).astype(np.int64)
        x_index_out, rank_out, send_counts = cls.construct_send_tables(comm, rank_in, x_index_in, recv_counts)
        A_i_nonlocal = cls._compact_columns(A_i, start_col, end_col, nonlocal_cols)
        comm.Barrier()
        return cls(A_i_local, A_i_nonlocal, x_index_in, rank_in, x_index_out, rank_out, all_n_i, start_col, end_col,
                   send_counts, recv_counts)

    @staticmethod
    def _compact_columns(A_i: sparse.csr_matrix, start_col: int, end_col: int, nonlocal_cols: np.ndarray) -> sparse.csr_matrix:
        """``A_i[:, nonlocal_cols]`` without SciPy's fancy column indexing (one pass over the entries)."""
        keep = ((A_i.indices < start_col) | (A_i.indices >= end_col)) & (A_i.data != 0)
        row_of = np.repeat(np.arange(A_i.shape[0], dtype=np.int64), np.diff(A_i.indptr))
        new_cols = np.searchsorted(nonlocal_cols, A_i.indices[keep])
        out = sparse.csr_matrix((A_i.data[keep], (row_of[keep], new_cols)), shape=(A_i.shape[0], nonlocal_cols.size),
                                dtype=A_i.dtype)
        out.sum_duplicates()
        out.sort_indices()
        return out
