"""Arrow decomposition of a sparse symmetric matrix without igraph (SURVEY.md "next" row N3).

Restates the algorithm of the reference's ``arrow/decomposition.py:32-281`` on scipy.sparse.csgraph:

* linear order of one level (``_arrow_linear_order``, ``:253-281``): the ``arrow_width`` highest-degree vertices
  first (they become the arrow head), then the vertices of positive degree arranged by a random spanning forest
  (``linearize_with_random_forest``, ``:165-205``: random edge weights -> minimum spanning forest -> rooted at the
  first vertex of each component -> preorder that visits smaller subtrees first, ``:208-240``; components of at
  most ``base_size`` vertices are appended as they are), then the isolated vertices;
* edges whose end points land within ``arrow_width`` of each other (band) or in the same block (block diagonal),
  or that touch the head, stay in this level (``:83-98``); the rest recurse into the next level (``:100-112``);
* the last allowed level takes everything that is left with a deterministic BFS order (``linearize_with_ck``,
  ``:147-162``) and reports the width it actually needs (``get_arrow_width``, ``:57-63``).

The result is a list of ``(B_j, permutation_j)`` exactly in the form ``graphio.save_decomposition_new`` writes and the
engine reads: ``B_j`` is the level's matrix in its own vertex order, ``permutation_j[r]`` the original vertex at row
``r``.  The random forest makes the output non-unique (the reference draws from ``numpy.random`` as well), so parity
is defined by the reference's own test properties (``tests/test_arrowdecomposition.py:24-112``): the levels partition
the edges, the permutations are permutations, non-final levels respect the width, and
``sum_j P_j B_j P_j^T == A``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
from scipy import sparse
from scipy.sparse import csgraph


def _forest_preorder(n: int, rows: np.ndarray, cols: np.ndarray, rng: np.random.Generator, base_size: int) -> np.ndarray:
    """Linear arrangement of a graph with ``n`` vertices of positive degree (edges ``rows[i] -- cols[i]``)."""
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    w = rng.random(rows.size) + 1e-9                         # random weights -> a random spanning forest
    G = sparse.coo_matrix((w, (rows, cols)), shape=(n, n)).tocsr()
    G = G.maximum(G.T)
    forest = csgraph.minimum_spanning_tree(G)
    forest = (forest + forest.T).tocsr()
    n_comp, label = csgraph.connected_components(forest, directed=False)
    comp_size = np.bincount(label, minlength=n_comp)
    # the reference roots every component at its first vertex (vertex 0 of the component subgraph)
    first = np.full(n_comp, n, dtype=np.int64)
    np.minimum.at(first, label, np.arange(n, dtype=np.int64))
    # one BFS from a virtual super-root attached to every component root gives parents + a topological order
    sup = n
    aug = sparse.vstack([sparse.hstack([forest, sparse.csr_matrix((n, 1))]),
                         sparse.csr_matrix((np.ones(n_comp), (np.zeros(n_comp, dtype=np.int64), first)), shape=(1, n + 1))]).tocsr()
    aug = aug.maximum(aug.T)
    order, pred = csgraph.breadth_first_order(aug, sup, directed=False, return_predecessors=True)
    order = order[1:]                                        # drop the super-root
    parent = pred[:n].astype(np.int64)
    parent[parent == sup] = -1
    # depth of every vertex by pointer doubling, then the vertices grouped by depth (everything below is one vectorised
    # pass per tree level instead of a Python loop per vertex / per component: a pruned power-law graph leaves 1e5-1e6
    # tiny components per level)
    depth = (parent >= 0).astype(np.int64)
    anc = parent.copy()
    while True:
        live = anc >= 0
        up = np.where(live, anc, 0)
        more = live & (parent[up] >= 0)
        if not more.any():
            break
        depth = np.where(live, depth + depth[up], depth)
        anc = np.where(live, anc[up], -1)
        # after a jump ``anc`` may land on a root: its depth contribution is already counted
    by_depth = np.argsort(depth, kind="stable")
    d_sorted = depth[by_depth]
    max_d = int(d_sorted[-1]) if n else 0
    cut = np.searchsorted(d_sorted, np.arange(max_d + 2), side="left")
    size = np.ones(n, dtype=np.int64)
    for d in range(max_d, 0, -1):                            # children before parents
        vs = by_depth[cut[d]:cut[d + 1]]
        np.add.at(size, parent[vs], size[vs])
    # children of every vertex sorted largest subtree first; the reference pushes them in that order and pops the
    # smallest first, so a vertex is visited after the siblings that FOLLOW it in this list: its pre-order offset below
    # its parent is 1 + the sizes of those siblings
    has_p = parent >= 0
    kids = np.flatnonzero(has_p)
    key = np.lexsort((-size[kids], parent[kids]))
    kids = kids[key]
    off = np.zeros(n, dtype=np.int64)
    if kids.size:
        ksz = size[kids]
        csum = np.cumsum(ksz)
        grp_first = np.flatnonzero(np.concatenate([[True], parent[kids][1:] != parent[kids][:-1]]))
        grp_id = np.cumsum(np.concatenate([[True], parent[kids][1:] != parent[kids][:-1]])) - 1
        grp_last = np.concatenate([grp_first[1:], [kids.size]]) - 1
        total_to_end = csum[grp_last][grp_id]                # cumulative size up to the end of the kid's sibling group
        off[kids] = 1 + (total_to_end - csum)                # sizes of the siblings listed after it
    # components in label order, like igraph's clustering; small ones keep ascending vertex ids, large ones are laid out
    # in the pre-order computed from the offsets, level by level
    comp_order = np.argsort(label, kind="stable")            # members of component c: ascending ids, contiguous
    comp_start = np.concatenate([[0], np.cumsum(comp_size)]).astype(np.int64)
    out = np.empty(n, dtype=np.int64)
    small = comp_size[label] <= base_size
    pre = np.full(n, -1, dtype=np.int64)
    # position of a small component's member = start of the component + its rank inside the component
    rank_in_comp = np.empty(n, dtype=np.int64)
    rank_in_comp[comp_order] = np.arange(n, dtype=np.int64) - comp_start[label[comp_order]]
    pre[small] = comp_start[label[small]] + rank_in_comp[small]
    roots = np.flatnonzero(~has_p & ~small)
    pre[roots] = comp_start[label[roots]]
    for d in range(1, max_d + 1):                            # parents before children
        vs = by_depth[cut[d]:cut[d + 1]]
        vs = vs[~small[vs]]
        pre[vs] = pre[parent[vs]] + off[vs]
    assert pre.min() >= 0
    out[pre] = np.arange(n, dtype=np.int64)
    return out


def _bfs_order(n: int, rows: np.ndarray, cols: np.ndarray, base_size: int = 2) -> np.ndarray:
    """Deterministic order of the last level: BFS per connected component (``linearize_with_ck``)."""
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    G = sparse.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(n, n)).tocsr()
    G = G.maximum(G.T)
    n_comp, label = csgraph.connected_components(G, directed=False)
    # components in the order of their first vertex; members grouped once (ascending ids inside a component)
    comp_size = np.bincount(label, minlength=n_comp)
    comp_order = np.argsort(label, kind="stable")
    comp_start = np.concatenate([[0], np.cumsum(comp_size)]).astype(np.int64)
    first = comp_order[comp_start[:-1]]
    # ONE breadth-first search from a virtual super-root attached to the first vertex of every large component: the
    # queue interleaves the components but keeps each component's own FIFO order, so the sub-sequence of a component is
    # exactly its stand-alone BFS order (a search per component re-validates the whole graph every time: minutes at 1e5
    # components)
    big = np.flatnonzero(comp_size > base_size)
    key = np.empty(n, dtype=np.int64)                        # position inside the component's block of the output
    rank_in_comp = np.empty(n, dtype=np.int64)
    rank_in_comp[comp_order] = np.arange(n, dtype=np.int64) - comp_start[label[comp_order]]
    key[:] = rank_in_comp                                    # small components: ascending vertex ids
    if big.size:
        sup = n
        roots = np.sort(first[big])
        aug = sparse.vstack([sparse.hstack([G, sparse.csr_matrix((n, 1))]),
                             sparse.csr_matrix((np.ones(roots.size), (np.zeros(roots.size, dtype=np.int64), roots)),
                                               shape=(1, n + 1))]).tocsr()
        aug = aug.maximum(aug.T)
        order = csgraph.breadth_first_order(aug, sup, directed=False, return_predecessors=False)[1:]
        key[order] = np.arange(order.size, dtype=np.int64)   # global BFS position: monotone inside every component
    comp_rank = np.empty(n_comp, dtype=np.int64)             # components in the order of their first vertex
    comp_rank[np.argsort(first, kind="stable")] = np.arange(n_comp, dtype=np.int64)
    return np.lexsort((key, comp_rank[label])).astype(np.int64)


def _linear_order(A: sparse.csr_matrix, arrow_width: int, deterministic: bool, rng: np.random.Generator) -> np.ndarray:
    n = A.shape[0]
    deg = np.diff(A.indptr)
    by_degree = np.argsort(-deg, kind="stable")              # highest degree first (ties: lowest id first)
    head = by_degree[:arrow_width]
    rest = by_degree[arrow_width:]
    middle = np.sort(rest[deg[rest] > 0])                    # igraph's induced subgraph numbers vertices by ascending id
    singles = rest[deg[rest] == 0]
    # sub-graph induced by the middle vertices, relabelled 0..m-1 in ascending original id
    m = middle.size
    relabel = np.full(n, -1, dtype=np.int64)
    relabel[middle] = np.arange(m)
    C = A.tocoo()
    keep = (relabel[C.row] >= 0) & (relabel[C.col] >= 0) & (C.row < C.col)
    r, c = relabel[C.row[keep]], relabel[C.col[keep]]
    if deterministic:
        sub = _bfs_order(m, r, c)
    else:
        sub = _forest_preorder(m, r, c, rng, min(arrow_width - 1, 16))
    order = np.concatenate([head, middle[sub], singles]).astype(np.int64)
    assert order.size == n
    return order


def arrow_decomposition(A, arrow_width: int = 512, max_number_of_levels: int = 2, block_diagonal: bool = False,
                        prune: bool = True, seed: Optional[int] = None, dtype=np.float32
                        ) -> List[Tuple[sparse.csr_matrix, np.ndarray]]:
    """Decompose the symmetric sparse matrix ``A`` (an undirected graph's adjacency, values kept) into arrow levels.

    Returns ``[(B_0, perm_0), (B_1, perm_1), ...]``; also records the width each level really needs in
    ``B_j.arrow_width`` (only the last, best-effort level can exceed ``arrow_width``)."""
    A = sparse.csr_matrix(A).astype(dtype)
    n = A.shape[0]
    assert A.shape[0] == A.shape[1] and arrow_width <= n
    A.sum_duplicates()
    A.sort_indices()
    rng = np.random.default_rng(seed)
    out: List[Tuple[sparse.csr_matrix, np.ndarray]] = []
    rest = A
    while True:
        last = len(out) + 1 >= max_number_of_levels
        order = _linear_order(rest, arrow_width, last, rng)
        inv = np.empty(n, dtype=np.int64)
        inv[order] = np.arange(n)
        C = rest.tocoo()
        pr, pc = inv[C.row], inv[C.col]
        if not last:
            if block_diagonal:
                near = (pr // arrow_width) == (pc // arrow_width)
            else:
                near = np.abs(pr - pc) <= arrow_width
            keep = near | ((pr < arrow_width) | (pc < arrow_width) if prune else False)
            if not np.any(keep):
                keep = np.ones_like(near)
        else:
            keep = np.ones(C.nnz, dtype=bool)
        B = sparse.csr_matrix((C.data[keep], (pr[keep], pc[keep])), shape=(n, n), dtype=dtype)
        B.sum_duplicates()
        B.sort_indices()
        width = arrow_width
        if last:
            far = (pr >= arrow_width) & (pc >= arrow_width)      # get_arrow_width (:57-63) without its off-by-one
            if np.any(far):
                width = max(width, int(np.max(np.abs(pr[far] - pc[far]))))
        B.arrow_width = width
        out.append((B, order))
        if np.all(keep):
            break
        rest = sparse.csr_matrix((C.data[~keep], (C.row[~keep], C.col[~keep])), shape=(n, n), dtype=dtype)
        rest.sum_duplicates()
        rest.sort_indices()
    return out


def reconstruct(decomposition, n: int) -> sparse.csr_matrix:
    """``sum_j P_j B_j P_j^T`` back in the original vertex order (what the reference's test subtracts from ``A``)."""
    total = sparse.csr_matrix((n, n), dtype=np.float64)
    for B, perm in decomposition:
        C = sparse.coo_matrix(B)
        total = total + sparse.csr_matrix((C.data.astype(np.float64), (perm[C.row], perm[C.col])), shape=(n, n))
    return total
