"""Helpers with the reference's names (``arrow/common/utils.py``), so driver scripts written against it keep working."""
from __future__ import annotations

from typing import Dict, Union

import numpy as np
from scipy import sparse

from .cli import str2bool  # noqa: F401  (utils.py:9-17)
from .synth import generate_dense_matrix, generate_sparse_matrix  # noqa: F401  (utils.py:63-99)


def time_to_ms(runtime: float) -> int:
    return int(runtime * 1000)


def mpi_print(rank: int, msg: str) -> None:
    """Print on rank 0 only (``utils.py:58-60``)."""
    if rank == 0:
        print(msg, flush=True)


def relabel_nodes(g: Union[sparse.csr_array, sparse.csr_matrix], mapping: Dict[int, int]):
    """``g'[mapping[a], mapping[b]] = g[a, b]`` (``utils.py:20-51``), by relabelling the COO coordinates instead of
    two sparse products with a permutation matrix."""
    if not isinstance(g, (sparse.csr_array, sparse.csr_matrix)):
        raise TypeError("The graph must be a SciPy-compatible CSR array or matrix.")
    if g.shape[0] != g.shape[1]:
        raise ValueError("The matrix must be square.")
    n = g.shape[0]
    labels = list(range(n))
    if sorted(mapping.keys()) != labels:
        raise ValueError("The keys of the mapping must be the rows of the graph's matrix representation.")
    if sorted(mapping.values()) != labels:
        raise ValueError("The values of the mapping must be the rows of the graph's matrix representation.")
    order = np.fromiter((mapping[i] for i in range(n)), dtype=np.int64, count=n)
    c = g.tocoo()
    out = type(g)((c.data, (order[c.row], order[c.col])), shape=g.shape)
    out.sum_duplicates()
    out.sort_indices()
    return out
