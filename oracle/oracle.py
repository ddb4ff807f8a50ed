"""ORACLE -- test infrastructure only (never imported by the product package).

CPU restatement of the reference's iterated arrow-decomposed SpMM hot path
(spcl/arrow-matrix @ a1965fa).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; the product
(`arrow_matrix_b200`) fails loudly without its CUDA library and never routes through here.

Parity pinning: the restatement is checked (tests/test_oracle.py, tests/test_golden.py) against
(1) the known-answer routing tables asserted by the reference's own ``tests/test_arrowmpi.py:24-94``,
(2) SciPy's ``csr_matrix @ ndarray`` (the reference's actual arithmetic) bit-for-bit, and
(3) golden vectors produced by running the UNMODIFIED reference classes in this container under
an in-process MPI stand-in (``tests/golden/make_golden.py``).

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy import sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    """liboracle.so (built by oracle/Makefile; __graft_entry__.build() compiles it too)."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(path)
    return _LIB


# --------------------------------------------------------------------------------------
# local arithmetic: scipy csr_matvecs  (arrow_slim_mpi.py:109-111, 125-127, 142-144)
# --------------------------------------------------------------------------------------
def csr_spmm_c(A: sparse.csr_matrix, X: np.ndarray, out: Optional[np.ndarray] = None,
               accumulate: bool = False) -> np.ndarray:
    """``A @ X`` through oracle/csr_matvecs.c (the restated SciPy kernel), fp32."""
    A = sparse.csr_matrix(A)
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, k = A.shape[0], X.shape[1]
    assert X.shape[0] == A.shape[1]
    if out is None:
        out = np.zeros((n, k), dtype=np.float32)
    elif not accumulate:
        out[:] = 0
    data = np.ascontiguousarray(A.data, dtype=np.float32)
    P = ctypes.c_void_p
    if A.indices.dtype == np.int64 or A.indptr.dtype == np.int64:
        ip = np.ascontiguousarray(A.indptr, dtype=np.int64)
        ix = np.ascontiguousarray(A.indices, dtype=np.int64)
        fn = _lib().oracle_csr_matvecs_f32_i64
    else:
        ip = np.ascontiguousarray(A.indptr, dtype=np.int32)
        ix = np.ascontiguousarray(A.indices, dtype=np.int32)
        fn = _lib().oracle_csr_matvecs_f32_i32
    fn.restype = None
    fn(ctypes.c_int64(n), ctypes.c_int64(k), P(ip.ctypes.data), P(ix.ctypes.data),
       P(data.ctypes.data), P(X.ctypes.data), P(out.ctypes.data))
    return out


def csr_spmm_scipy(A: sparse.csr_matrix, X: np.ndarray) -> np.ndarray:
    """The reference's literal call: ``A @ X``."""
    return A @ X


# --------------------------------------------------------------------------------------
# global view: tests/test_arrowdecomposition.py:139-156 (compute_spmm)
# --------------------------------------------------------------------------------------
def compute_spmm(decomposition: Sequence[Tuple[sparse.csr_matrix, np.ndarray]], X: np.ndarray) -> np.ndarray:
    """``sum_j (B_j @ X[perm_j])[argsort(perm_j)]`` in the precision of ``X`` (fp32 = the reference tests' own golden;
    float64 inputs give the exact yardstick)."""
    acc = np.zeros((X.shape[0], X.shape[1]), dtype=X.dtype)
    for B, perm in decomposition:
        inv = np.argsort(perm)
        acc += (B @ X[perm])[inv]
    return acc


# --------------------------------------------------------------------------------------
# loader semantics: arrow_dec_mpi.py:612-627, 695-749 ; graphio.py:361-406
# --------------------------------------------------------------------------------------
def number_of_blocks(adjacency, width: int) -> int:
    """ceil(#rows up to the last non-empty row / width)  (arrow_dec_mpi.py:612-627)."""
    if isinstance(adjacency, tuple):
        indptr = np.asarray(adjacency[2])
    else:
        indptr = sparse.csr_matrix(adjacency).indptr
    per_row = np.diff(indptr)
    nz = np.flatnonzero(per_row > 0)
    if nz.size == 0:
        raise StopIteration("matrix has no non-zero row")  # the reference's next() raises here too
    return int(-(-(int(nz[-1]) + 1) // width))


def prepare_permutations(perms: Sequence[np.ndarray], n_blocks: Sequence[int], width: int):
    """One-based fix-up, identity padding to ``n_blocks[0]*width`` and the to_prev / to_next maps.

    Follows arrow_dec_mpi.py:699-749: ``one_based = min(perm_0) > 0``; every permutation is padded
    with ``arange(old, rows)``; ``to_prev_j = inv_{j-1}[perm_j]``, ``to_next_j = inv_{j+1}[perm_j]``;
    entries ``>= width*n_blocks[neighbour]`` become the sentinel ``2*width*n_blocks[0]``.
    """
    rows = int(n_blocks[0]) * width
    sentinel = 2 * width * int(n_blocks[0])
    perms = [np.array(p, dtype=np.int64, copy=True) for p in perms]
    one_based = bool(np.min(perms[0]) > 0)
    for i in range(len(perms)):
        if one_based:
            perms[i] -= 1
        if perms[i].size < rows:
            perms[i] = np.concatenate([perms[i], np.arange(perms[i].size, rows, dtype=np.int64)])
        assert perms[i].size == rows, "permutation longer than n_blocks[0]*width (reference asserts, :714)"
    inv = [np.argsort(p) for p in perms]
    L = len(perms)
    to_prev: List[Optional[np.ndarray]] = [None] * L
    to_next: List[Optional[np.ndarray]] = [None] * L
    for i in range(L):
        if i > 0:
            t = inv[i - 1][perms[i]]
            to_prev[i] = np.where(t >= width * int(n_blocks[i - 1]), sentinel, t).astype(np.int64)
        if i < L - 1:
            t = inv[i + 1][perms[i]]
            to_next[i] = np.where(t >= width * int(n_blocks[i + 1]), sentinel, t).astype(np.int64)
    return perms, to_prev, to_next, sentinel


def arrow_mask(B: sparse.csr_matrix, width: int, n_blocks: int, block_diagonal: bool = True) -> sparse.csr_matrix:
    """The part of level ``B`` the reference actually multiplies with.

    ``split_matrix_to_blocks`` (graphio.py:382-383) keeps blocks (0,j), (i,0), (i,i) and -- banded
    mode -- (i,i+-1); the loader truncates to ``n_blocks`` block-rows/columns
    (arrow_dec_mpi.py:728-731).  Everything else is dropped silently.
    """
    n = n_blocks * width
    C = sparse.coo_matrix(sparse.csr_matrix(B))
    bi, bj = C.row // width, C.col // width
    keep = (C.row < n) & (C.col < n)
    pat = (bi == 0) | (bj == 0) | (bi == bj)
    if not block_diagonal:
        pat |= (np.abs(bi - bj) == 1)
    keep &= pat
    M = sparse.csr_matrix((C.data[keep], (C.row[keep], C.col[keep])), shape=(n, n), dtype=np.float32)
    M.sum_duplicates()
    M.sort_indices()
    return M


# --------------------------------------------------------------------------------------
# routing tables: arrow_dec_mpi.py:325-384
# --------------------------------------------------------------------------------------
def all_to_all_tables(out_permutation: np.ndarray, rows_per_rank: int, n_columns: int,
                      total_ranks: int, put_offset: int = 0):
    """counts, displs, pack order, unpack order -- same four outputs as ``_all_to_all_tables``."""
    out_permutation = np.asarray(out_permutation)
    assert out_permutation.size == rows_per_rank and put_offset < total_ranks
    assert n_columns > 0 and total_ranks > 0
    ranks = (out_permutation // rows_per_rank).astype(np.intp)
    ok = ranks + put_offset < total_ranks                                    # :351
    counts = np.bincount((ranks[ok] + put_offset).astype(np.intp), minlength=total_ranks).astype(np.int64) * n_columns
    displs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)  # :356-359
    send_perm = np.argsort(ranks, kind="stable")                             # :363
    sel = np.flatnonzero(ranks < total_ranks)                                # :376 (no put_offset here)
    order = np.lexsort((out_permutation[sel], ranks[sel]))                   # group by rank, then by target row
    recv_perm = sel[order].astype(np.intp)
    return [int(c) for c in counts], [int(d) for d in displs], send_perm, recv_perm


# --------------------------------------------------------------------------------------
# the per-iteration protocol: arrow_dec_mpi.py:283-307 (step), 404-440, 507-550;
# arrow_slim_mpi.py:104-155 (block algebra)
# --------------------------------------------------------------------------------------
class ReferenceProtocolOracle:
    """Global-array restatement of what the reference's ranks jointly compute in ``step()``.

    Per level ``j`` it keeps the concatenation of all column ranks' tiles.  The aliasing the
    reference relies on is reproduced: after every exchange ``X`` *is* ``C``
    (arrow_dec_mpi.py:438, 545), so rows whose ``to_prev`` is the sentinel keep the previous
    iteration's result (``:544`` only overwrites routed rows), and the SpMM rebinds ``C`` to a
    fresh array (arrow_slim_mpi.py:125-127).
    """

    def __init__(self, decomposition: Sequence[Tuple[sparse.csr_matrix, np.ndarray]], width: int,
                 k: int, block_diagonal: bool = True, n_blocks: Optional[Sequence[int]] = None,
                 use_c_kernel: bool = False, blockwise: bool = False, dtype=np.float32):
        """``dtype=np.float64`` turns the restatement into the exact-arithmetic yardstick the parity tests use to
        tell rounding (any fp32 summation order, the reference's included) from a wrong result; the reference itself
        computes in fp32 (arrow_bench.py:21)."""
        self.dtype = np.dtype(dtype)
        self.width, self.k = width, k
        self.L = len(decomposition)
        self.n_blocks = [number_of_blocks(B, width) for B, _ in decomposition] if n_blocks is None else list(n_blocks)
        self.perms, self.to_prev, self.to_next, self.sentinel = prepare_permutations(
            [p for _, p in decomposition], self.n_blocks, width)
        self.rows = [nb * width for nb in self.n_blocks]
        self.mats = [arrow_mask(B, width, nb, block_diagonal) for (B, _), nb in zip(decomposition, self.n_blocks)]
        self.dropped_nnz = [int(sparse.csr_matrix(B).nnz - M.nnz) for (B, _), M in zip(decomposition, self.mats)]
        if self.dtype != np.float32:
            assert not use_c_kernel, "the C restatement of csr_matvecs is fp32 like SciPy's instantiation the reference uses"
            self.mats = [M.astype(self.dtype) for M in self.mats]
        self.C = [np.zeros((r, k), dtype=self.dtype) for r in self.rows]       # zero_rhs (arrow_slim_mpi.py:354-394)
        self.X = [np.zeros((r, k), dtype=self.dtype) for r in self.rows]
        self._mm = csr_spmm_c if use_c_kernel else (lambda A, X: A @ X)
        self.blockwise = blockwise
        self.block_diagonal = block_diagonal

    def set_features(self, X0: np.ndarray) -> None:
        """Level-0 tiles, in level-0 (permuted) row order; stored by reference like ``set_features``."""
        assert X0.shape == (self.rows[0], self.k)
        self.X[0] = X0 if X0.dtype == self.dtype else X0.astype(self.dtype)

    # forward exchange, arrow_dec_mpi.py:507-550
    def propagate_features(self) -> None:
        for j in range(1, self.L):
            tp = self.to_prev[j][: self.rows[j]]
            ok = tp < self.rows[j - 1]
            self.C[j][ok] = self.X[j - 1][tp[ok]]       # :544  C_i[back_receive_permutation] = recvbuf
            self.X[j] = self.C[j]                       # :545  set_features(C_i)  (alias)

    def _spmm_level(self, j: int) -> np.ndarray:
        if not self.blockwise:
            return np.asarray(self._mm(self.mats[j], self.X[j]), dtype=self.dtype)
        # block algebra of arrow_slim_mpi.py:104-155: C_0 = sum_i A_0i X_i ; C_i = A_ii X_i + A_i0 X_0
        w, t = self.width, self.n_blocks[j]
        M, X = self.mats[j], self.X[j]
        out = np.zeros_like(X)
        X0 = X[:w]
        c0 = np.zeros((w, self.k), dtype=self.dtype)
        for i in range(t):
            c0 += M[:w, i * w:(i + 1) * w] @ X[i * w:(i + 1) * w]
        out[:w] = c0
        for i in range(1, t):
            ci = M[i * w:(i + 1) * w, i * w:(i + 1) * w] @ X[i * w:(i + 1) * w]
            ci += M[i * w:(i + 1) * w, :w] @ X0
            if not self.block_diagonal:                     # banded layout, arrow_mpi.py:211-219
                if i > 1:
                    ci += M[i * w:(i + 1) * w, (i - 1) * w:i * w] @ X[(i - 1) * w:i * w]
                if i < t - 1:
                    ci += M[i * w:(i + 1) * w, (i + 1) * w:(i + 2) * w] @ X[(i + 1) * w:(i + 2) * w]
            out[i * w:(i + 1) * w] = ci
        return out

    def spmm(self) -> None:
        for j in range(self.L):
            self.C[j] = self._spmm_level(j)             # fresh array; X[j] keeps the old object

    # backward exchange, arrow_dec_mpi.py:404-440
    def aggregate(self) -> None:
        for j in range(self.L - 1, 0, -1):
            tp = self.to_prev[j][: self.rows[j]]
            ok = tp < self.rows[j - 1]
            self.C[j - 1][tp[ok]] += self.C[j][ok]      # :437 (to_prev is injective on routed rows)
            self.X[j - 1] = self.C[j - 1]               # :438

    def step(self) -> np.ndarray:
        """One ``ArrowDecompositionMPI.step()``; returns level-0 ``C`` (which is also the new ``X``)."""
        self.propagate_features()
        self.spmm()
        self.aggregate()
        return self.C[0]


def to_original_order(C0: np.ndarray, perm0: np.ndarray, n: int) -> np.ndarray:
    """Level-0 row order -> original vertex order (inverse of ``X[perm_0]``, test_arrowmpi.py:275, 290)."""
    out = np.zeros((n, C0.shape[1]), dtype=C0.dtype)
    m = min(n, perm0.size, C0.shape[0])
    valid = perm0[:m] < n
    out[perm0[:m][valid]] = C0[:m][valid]
    return out
