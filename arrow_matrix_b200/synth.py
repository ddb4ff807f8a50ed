"""Synthetic inputs for the arrow SpMM path (no igraph).

The reference synthesises a Barabasi-Albert graph and decomposes it with igraph
(``arrow/arrow_bench.py:28-41``); igraph is not available here, so the benchmark inputs are
generated directly in decomposed form:

* ``generate_sparse_matrix`` / ``generate_dense_matrix`` follow the recipes (and value
  distributions) of reference ``arrow/common/utils.py:63-99``.
* ``arrow_csr`` builds an arrow-shaped level: rows of block-row 0 reach every column, every
  other row has ``head_nnz`` entries in the head columns ``[0, w)`` and ``diag_nnz`` in its own
  diagonal block.  Columns are drawn one per stratum, so rows come out sorted and duplicate
  free (canonical CSR, which the reference asserts at ``arrow_slim_mpi.py:306-308``).
* ``synth_decomposition`` chains levels with permutations (identity / random / shard-local).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
from scipy import sparse


def generate_sparse_matrix(rows: int, cols: int, nnz: int, dtype, rng: np.random.Generator) -> sparse.csr_matrix:
    """Fixed draws per row, uniform columns, then canonicalised (reference ``utils.py:63-87``)."""
    per_row = int(np.ceil(nnz / rows))
    total = per_row * rows
    data = rng.random((total,), dtype=dtype)
    indptr = np.arange(0, total + 1, per_row, dtype=np.int64)
    indices = rng.integers(0, cols, size=(total,), dtype=np.int64)
    m = sparse.csr_matrix((data, indices, indptr), shape=(rows, cols), dtype=dtype)
    m.sum_duplicates()
    m.sort_indices()
    return m


def generate_dense_matrix(rows: int, cols: int, dtype, rng: np.random.Generator) -> np.ndarray:
    """U[-1, 1) features (reference ``utils.py:90-99``; same as ``arrow_bench.py:115``)."""
    return 2 * rng.random((rows, cols), dtype=dtype) - 1


def _stratified(rng: np.random.Generator, n_rows: int, draws: int, lo: np.ndarray, span: int) -> np.ndarray:
    """``draws`` sorted distinct columns per row inside ``[lo, lo+span)`` (one per stratum)."""
    edges = (np.arange(draws + 1, dtype=np.int64) * span) // draws
    width = np.diff(edges)
    if np.any(width <= 0):
        raise ValueError("span too small for the requested draws per row")
    u = rng.random((n_rows, draws), dtype=np.float32)
    off = np.minimum((u * width[None, :]).astype(np.int64), width[None, :] - 1)
    return lo[:, None] + edges[None, :-1] + off


def arrow_csr(n_total: int, width: int, n_active_blocks: int, rng: np.random.Generator,
              head_nnz: int = 3, diag_nnz: int = 7, dtype=np.float32,
              hub_rows: int = 0, hub_nnz: int = 0, band_nnz: int = 0) -> sparse.csr_matrix:
    """Arrow-shaped ``n_total x n_total`` CSR whose first ``n_active_blocks*width`` rows are non-empty.

    Block-row 0 spreads ``head_nnz+diag_nnz`` entries over all active columns; block-row ``i>0``
    puts ``head_nnz`` entries in ``[0, w)`` and ``diag_nnz`` in ``[i*w, (i+1)*w)``.  ``hub_rows``
    of the head get ``hub_nnz`` entries instead (degree skew like a real decomposition's hubs).
    """
    n_act = n_active_blocks * width
    assert n_act <= n_total
    head_rows = min(width, n_act)
    per_head = min(head_nnz + diag_nnz, n_act)           # block-row 0: spread over every active column
    h = min(head_nnz, width)
    d = min(diag_nnz, width)
    per_rest = h + d
    rest = n_act - head_rows
    head_cols = _stratified(rng, head_rows, per_head, np.zeros(head_rows, dtype=np.int64), n_act)
    pieces = [head_cols.reshape(-1)]
    counts = np.zeros(n_total, dtype=np.int64)
    counts[:head_rows] = per_head
    if rest > 0:
        rc = np.empty((rest, per_rest), dtype=np.int64)
        rc[:, :h] = _stratified(rng, rest, h, np.zeros(rest, dtype=np.int64), width)
        blk_lo = (np.arange(width, n_act, dtype=np.int64) // width) * width
        rc[:, h:] = _stratified(rng, rest, d, blk_lo, width)
        pieces.append(rc.reshape(-1))
        counts[head_rows:n_act] = per_rest
    indices = np.concatenate(pieces)
    data = rng.random((indices.size,), dtype=dtype)
    if hub_rows > 0 and hub_nnz > per_head:
        # replace the first hub_rows rows by wide rows
        hub_rows = min(hub_rows, head_rows)
        hub_nnz = min(hub_nnz, n_act)
        hub_cols = _stratified(rng, hub_rows, hub_nnz, np.zeros(hub_rows, dtype=np.int64), n_act)
        hub_data = rng.random((hub_rows * hub_nnz,), dtype=dtype)
        indices = np.concatenate([hub_cols.reshape(-1), indices[hub_rows * per_head:]])
        data = np.concatenate([hub_data, data[hub_rows * per_head:]])
        counts[:hub_rows] = hub_nnz
    indptr = np.zeros(n_total + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    idx_dtype = np.int32 if n_total < 2**31 and indptr[-1] < 2**31 else np.int64
    m = sparse.csr_matrix((data, indices.astype(idx_dtype), indptr.astype(idx_dtype)),
                          shape=(n_total, n_total), dtype=dtype)
    m.has_sorted_indices = True
    m.has_canonical_format = True
    if band_nnz > 0 and n_active_blocks > 2:
        # banded (non block-diagonal) variant: extra entries in the neighbouring blocks (i, i-1) and (i, i+1)
        # for block-rows i >= 1 (the reference's arrow-banded shape, arrow_mpi.py:211-219)
        rows = np.repeat(np.arange(width, n_act, dtype=np.int64), band_nnz)
        bi = rows // width
        side = rng.integers(0, 2, size=rows.size) * 2 - 1
        bj = bi + side
        bj = np.where(bj < 1, bi + 1, bj)
        bj = np.where(bj >= n_active_blocks, bi - 1, bj)
        ok = (bj >= 1) & (bj < n_active_blocks) & (bj != bi)
        cols = bj * width + rng.integers(0, width, size=rows.size)
        extra = sparse.csr_matrix((rng.random(int(ok.sum()), dtype=dtype), (rows[ok], cols[ok])),
                                  shape=(n_total, n_total), dtype=dtype)
        m = sparse.csr_matrix(m + extra)
        m.sum_duplicates()
        m.sort_indices()
    return m


def make_permutation(n: int, kind: str, rng: np.random.Generator, shards: int = 8) -> np.ndarray:
    """Level permutation: 'identity', 'random' (worst-case exchange) or 'local' (shuffled inside n/shards slabs)."""
    if kind == "identity":
        return np.arange(n, dtype=np.int64)
    if kind == "random":
        return rng.permutation(n).astype(np.int64)
    if kind == "local":
        p = np.arange(n, dtype=np.int64)
        bounds = (np.arange(shards + 1, dtype=np.int64) * n) // shards
        for s in range(shards):
            seg = p[bounds[s]:bounds[s + 1]]
            rng.shuffle(seg)
        return p
    raise ValueError(f"unknown permutation kind {kind!r}")


def synth_decomposition(n_blocks0: int, width: int, levels: int = 2, perm_kind: str = "random",
                        seed: int = 503, shrink: int = 2, hub_rows: int = 0, hub_nnz: int = 0,
                        head_nnz: int = 3, diag_nnz: int = 7, nested: bool = True, band_nnz: int = 0,
                        ) -> List[Tuple[sparse.csr_matrix, np.ndarray]]:
    """G2 of SURVEY.md 8d: ``levels`` arrow matrices over ``n = n_blocks0*width`` vertices.

    Level 0 uses the identity permutation and fills every row; level ``j`` keeps
    ``n_blocks0 / shrink**j`` active block-rows (the rest are empty rows, like the zero-degree tail
    of a real decomposition) under a ``perm_kind`` permutation.

    ``nested=True`` keeps the vertices active at level ``j`` inside those active at level ``j-1``
    (true for real decompositions: level ``j`` only holds edges left over from level ``j-1``), so the
    reference's level-to-level feature chain never meets its sentinel.  ``nested=False`` draws the
    permutations independently and exercises the sentinel / stale-row behaviour
    (reference ``arrow_dec_mpi.py:740-749, 544``).
    """
    n = n_blocks0 * width
    rng = np.random.default_rng(seed)
    out = []
    for j in range(levels):
        act = max(1, n_blocks0 // (shrink ** j))
        mat = arrow_csr(n, width, act, rng, head_nnz=head_nnz, diag_nnz=diag_nnz,
                        hub_rows=hub_rows if j == 0 else 0, hub_nnz=hub_nnz, band_nnz=band_nnz)
        perm = make_permutation(n, "identity" if j == 0 else perm_kind, rng)
        if nested and j >= 2:
            # re-draw so that positions [0, prev_act) of this level stay inside the previous level's active rows
            prev_act = max(1, n_blocks0 // (shrink ** (j - 1))) * width
            q = np.concatenate([make_permutation(prev_act, perm_kind, rng),
                                prev_act + make_permutation(n - prev_act, perm_kind, rng)])
            perm = out[-1][1][q]
        out.append((mat, perm))
    return out


def barabasi_albert(n: int, m: int, seed: int = 503, dtype=np.float32) -> sparse.csr_matrix:
    """Adjacency matrix of a Barabasi-Albert preferential-attachment graph (the reference's synthetic input,
    ``igraph.Graph.Barabasi(n, m, 503)`` at ``arrow_bench.py:33``; igraph is unavailable, same model, different stream)."""
    rng = np.random.default_rng(seed)
    m = max(1, min(m, n - 1))
    src, dst = [], []
    pool = list(range(m))                       # start: m isolated vertices, each counted once
    for v in range(m, n):
        targets = set()
        while len(targets) < m:
            targets.add(pool[int(rng.integers(0, len(pool)))])
        for t in targets:
            src.append(v)
            dst.append(t)
        pool.extend(targets)
        pool.extend([v] * m)
    r = np.asarray(src + dst, dtype=np.int64)
    c = np.asarray(dst + src, dtype=np.int64)
    A = sparse.csr_matrix((np.ones(r.size, dtype=dtype), (r, c)), shape=(n, n))
    A.sum_duplicates()
    A.data[:] = 1
    A.sort_indices()
    return A
