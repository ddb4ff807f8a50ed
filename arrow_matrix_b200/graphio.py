"""On-disk layout of an arrow decomposition (read + write side, no igraph).

Drop-in for the read side of the reference's ``arrow/common/graphio.py``:

* file naming            -> reference ``graphio.py:38-70``  (``format_path``)
* npy triplet + perm     -> reference ``graphio.py:251-314`` (``load_decomposition_new``)
* arrow block splitting  -> reference ``graphio.py:361-406`` (``split_matrix_to_blocks``)

Level ``i`` of a decomposition of base name ``base`` and width ``w`` lives in

    {base}_B_{w}_{i}[_bd]_indptr.npy / _indices.npy / _data.npy (optional) / _permutation.npy

``_data.npy`` may be missing (the Julia converter never writes it,
``julia/arrow/convert_to_csr.jl:45-66``): values then default to 1.0f.  Each level is the
full ``n x n`` CSR of that level in its own permuted vertex order; ``permutation[r]`` is
the original vertex id sitting at row ``r``.
"""
from __future__ import annotations

import enum
import os
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
from scipy import sparse


class DecompositionFileType(enum.Enum):
    npz = 1
    indptr_npy = 2
    indices_npy = 3
    data_npy = 4
    permutation_npy = 5
    nonzero_rows_npy = 6


_SUFFIX = {
    DecompositionFileType.npz: ".npz",
    DecompositionFileType.indptr_npy: "_indptr.npy",
    DecompositionFileType.indices_npy: "_indices.npy",
    DecompositionFileType.data_npy: "_data.npy",
    DecompositionFileType.permutation_npy: "_permutation.npy",
    DecompositionFileType.nonzero_rows_npy: "_nnzrows.npy",
}


def format_path(base_path: str, width: int, index: Optional[int], block_diagonal: bool,
                file_type: DecompositionFileType) -> str:
    """Same string as reference ``graphio.py:38-70`` (note: ``_B`` is appended here)."""
    parts = [f"{base_path}_B", f"_{width}"]
    if index is not None:
        parts.append(f"_{index}")
    if block_diagonal:
        parts.append("_bd")
    parts.append(_SUFFIX[file_type])
    return "".join(parts)


def get_pathname(basename: str, width: int, is_block_diagonal: bool) -> str:
    """Display name used in log lines (reference ``graphio.py:498-504``)."""
    name = f"{basename}_B"
    if width:
        name += f"_{width}"
    if is_block_diagonal:
        name += "_bd"
    return name


def decomposition_size(filename: str, width: int, block_diagonal: bool) -> int:
    """Number of consecutive levels that have a permutation file (reference ``:120-128``)."""
    i = 0
    while os.path.exists(format_path(filename, width, i, block_diagonal,
                                     DecompositionFileType.permutation_npy)):
        i += 1
    return i


def save_decomposition_new(decomposition: Sequence[Tuple[sparse.csr_matrix, np.ndarray]],
                           filename: str, width: int, block_diagonal: bool = True,
                           write_data: bool = True, index_dtype=None,
                           one_based_permutation: bool = False) -> None:
    """Write levels in the npy layout the reference reads.

    Mirrors the write side of reference ``graphio.py:131-191`` without the igraph objects:
    the caller hands in ``(csr, permutation)`` pairs directly.  ``write_data=False``,
    ``index_dtype=np.int64`` and ``one_based_permutation=True`` reproduce the Julia
    converter's quirks so the loader can be tested against them.
    """
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    for i, (mat, perm) in enumerate(decomposition):
        mat = sparse.csr_matrix(mat)
        indptr, indices = mat.indptr, mat.indices
        if index_dtype is not None:
            indptr, indices = indptr.astype(index_dtype), indices.astype(index_dtype)
        np.save(format_path(filename, width, i, block_diagonal, DecompositionFileType.indptr_npy), indptr)
        np.save(format_path(filename, width, i, block_diagonal, DecompositionFileType.indices_npy), indices)
        if write_data:
            np.save(format_path(filename, width, i, block_diagonal, DecompositionFileType.data_npy), mat.data)
        perm = np.asarray(perm)
        if one_based_permutation:
            perm = perm + 1
        np.save(format_path(filename, width, i, block_diagonal, DecompositionFileType.permutation_npy), perm)


def save_decomposition(decomposition: Sequence[Tuple[sparse.csr_matrix, np.ndarray]], filename: str, width: int,
                       block_diagonal: bool = True, dtype=np.float32) -> None:
    """The reference's other layout (``graphio.py:73-117``): level ``i`` as one SciPy ``.npz`` next to its
    ``_permutation.npy``, plus the ``_nnzrows.npy`` convenience file (not read by the SpMM path)."""
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    nonzero_rows = []
    for i, (mat, perm) in enumerate(decomposition):
        mat = sparse.csr_matrix(mat).astype(dtype)
        sparse.save_npz(format_path(filename, width, i, block_diagonal, DecompositionFileType.npz), mat)
        np.save(format_path(filename, width, i, block_diagonal, DecompositionFileType.permutation_npy), np.asarray(perm))
        # the reference stores the number of isolated vertices under this name (decomposition.py:20)
        nonzero_rows.append(int(np.count_nonzero(np.diff(mat.indptr) == 0)))
    np.save(format_path(filename, width, 0, block_diagonal, DecompositionFileType.nonzero_rows_npy),
            np.asarray(nonzero_rows, dtype=np.int64))


def load_decomposition(filename: str, width: int = None, block_diagonal: bool = True, no_permutation: bool = False
                       ) -> List[Tuple[sparse.csr_matrix, Optional[np.ndarray]]]:
    """Read the ``.npz`` layout (``graphio.py:194-249``), including the reference's fallback to its OLD file naming
    (``{base}_B_{width}_{i}_bd.npz``: the level index before the ``_bd`` marker) when the current one finds nothing."""
    out = []
    for i in range(decomposition_size(filename, width, block_diagonal)):
        B = sparse.csr_matrix(sparse.load_npz(format_path(filename, width, i, block_diagonal, DecompositionFileType.npz)))
        perm = None if no_permutation else np.load(format_path(filename, width, i, block_diagonal,
                                                                DecompositionFileType.permutation_npy))
        out.append((B, perm))
    if not out:
        i = 0
        while True:
            base = f"{filename}_B" + (f"_{width}" if width else "") + f"_{i}" + ("_bd" if block_diagonal else "")
            if not os.path.exists(base + ".npz"):
                break
            B = sparse.csr_matrix(sparse.load_npz(base + ".npz"))
            perm = None if no_permutation else np.load(base + "_permutation.npy")
            out.append((B, perm))
            i += 1
    return out


CsrTriplet = Tuple[np.ndarray, np.ndarray, np.ndarray]  # (data, indices, indptr) like the reference's mmap tuple


def load_decomposition_new(filename: str, width: int = None, block_diagonal: bool = True,
                           no_permutation: bool = False, mem_map: bool = False
                           ) -> List[Tuple[Union[sparse.csr_matrix, CsrTriplet], Optional[np.ndarray]]]:
    """Read every level until a file is missing (reference ``graphio.py:251-314``).

    ``mem_map=True`` returns the raw ``(data, indices, indptr)`` memmaps like the reference's
    ``:299-300`` -- this is what the sharded device loader consumes.
    """
    out = []
    i = 0
    while True:
        try:
            p = format_path(filename, width, i, block_diagonal, DecompositionFileType.indptr_npy)
            indptr = np.lib.format.open_memmap(p, mode="r") if mem_map else np.load(p)
            p = format_path(filename, width, i, block_diagonal, DecompositionFileType.indices_npy)
            indices = np.lib.format.open_memmap(p, mode="r") if mem_map else np.load(p)
            p = format_path(filename, width, i, block_diagonal, DecompositionFileType.data_npy)
            if os.path.exists(p):
                data = np.lib.format.open_memmap(p, mode="r") if mem_map else np.load(p)
            else:
                # value-less level files (Julia converter): the reference hands out ones in both modes (graphio.py:292-298);
                # the memory-mapped route gets them as a zero-stride broadcast view: no nnz-sized allocation, and every
                # consumer of the documented (data, indices, indptr) triplet can slice it
                data = np.broadcast_to(np.float32(1), (indices.size,)) if mem_map else np.ones(indices.size, dtype=np.float32)
            if mem_map:
                B = (data, indices, indptr)
            else:
                # the reference lets SciPy infer the shape (graphio.py:302), which shrinks the column count when the
                # trailing vertices are isolated; levels are square adjacency matrices, so state that explicitly
                n = indptr.size - 1
                square = indices.size == 0 or int(indices.max()) < n
                B = sparse.csr_matrix((data, indices, indptr), shape=(n, n)) if square \
                    else sparse.csr_matrix((data, indices, indptr))
            perm = None
            if not no_permutation:
                perm = np.load(format_path(filename, width, i, block_diagonal,
                                           DecompositionFileType.permutation_npy))
        except FileNotFoundError:
            break
        out.append((B, perm))
        i += 1
    return out


def split_matrix_to_blocks(A: sparse.csr_matrix, block_size: int, dtype=None,
                           use_min_shape: bool = False) -> List[List[Optional[sparse.csr_matrix]]]:
    """Cut ``A`` into ``block_size`` squares, keeping only the arrow pattern.

    Behaviour of reference ``graphio.py:361-406``: block ``(i, j)`` is materialised only for
    ``i == 0`` or ``j in (0, i-1, i, i+1)``; a short last block-row is padded to ``block_size``
    rows by repeating the final ``indptr`` entry (its column count stays short unless the
    matrix is square and padded too -- the padded constructor forces ``block_size`` columns).
    Everything outside the pattern is dropped, silently, exactly as the reference does.
    """
    A = sparse.csr_matrix(A)
    rows, cols = A.shape
    dtype = dtype or A.dtype
    nbr = -(-rows // block_size)
    nbc = -(-cols // block_size)
    blocks: List[List[Optional[sparse.csr_matrix]]] = [[None] * nbc for _ in range(nbr)]
    for i in range(nbr):
        r0, r1 = i * block_size, min(rows, (i + 1) * block_size)
        row_slab = A[r0:r1, :]
        for j in range(nbc):
            if i > 0 and j not in (0, i - 1, i, i + 1):
                continue
            c0, c1 = j * block_size, min(cols, (j + 1) * block_size)
            piece = sparse.csr_matrix(row_slab[:, c0:c1])
            short = block_size - (r1 - r0)
            if use_min_shape or short == 0:
                blk = sparse.csr_matrix(piece, shape=(r1 - r0, c1 - c0), dtype=dtype)
            else:
                ip = np.concatenate([piece.indptr, np.full(short, piece.indptr[-1], dtype=piece.indptr.dtype)])
                blk = sparse.csr_matrix((piece.data, piece.indices, ip), shape=(block_size, block_size), dtype=dtype)
            blk.sum_duplicates()
            blk.sort_indices()
            blocks[i][j] = blk
    return blocks
