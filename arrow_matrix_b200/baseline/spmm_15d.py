"""A-stationary 1.5D SpMM with replication factor ``c`` on the library's kernels (SURVEY.md "next" row N4).

Mirror of the reference's ``arrow/baseline/spmm_15d.py``.  ``P`` ranks (GPUs) form a ``(P/c) x c`` grid, rank
``= x*c + y`` (row-major like ``Create_cart``, ``:43-47``); ``rounds = P/c^2``:

* rank ``(x, y)`` keeps the block ``A[x*lNI:(x+1)*lNI, y*lNK:(y+1)*lNK]`` with ``lNI = ceil(NI/(P/c))``,
  ``lNKb = ceil(NK/(P/c))``, ``lNK = lNKb*rounds`` (``:78-81, :87-96``), cut into ``rounds`` column blocks of width
  ``lNKb`` (``:120-127``);
* the rows ``[x*lNKb, (x+1)*lNKb)`` of ``X`` live on every rank of grid row ``x`` (``:134-152``);
* one product (``:313-367``): for ``r`` in rounds, the ranks of grid column ``y`` take ``X`` block ``q = y*rounds+r``
  from rank ``(q, y)`` and accumulate ``Y += A_r @ X_q``; then ``Y`` is all-reduced over the ``c`` ranks of a grid row.

On the device the broadcast is a peer copy of the owner's resident ``X`` tile over NVLink (sequential reads; no host
staging), the blocks are uploaded once (the reference re-uploads ``A_r`` in every round, ``:406``), and the all-reduce
is ``c`` peer reads summed in rank order, so every replica holds the same bits.  Barriers: X tiles final -> rounds ->
partials final -> reduction -> tiles reusable.

Reference-facing names are kept (``generate_15d_decomposition[_new]``, ``spmm_15d_gpu``); the three communicator
arguments are one ``Grid15D`` object.  ``spmm_15d_cpu`` raises -- this package has no CPU arithmetic.
"""
from __future__ import annotations

import math
import os
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy import sparse

from .. import comm as comm_mod
from .. import synth, wb_logging
from .spmm_petsc import CudaHaloFabric


class Grid15D:
    """The ``(P/c) x c`` process grid: stands in for the reference's ``cart_comm`` / ``bcast_comm`` / ``reduce_comm``."""

    def __init__(self, comm, c: int):
        self.comm = comm
        self.size, self.rank = comm.Get_size(), comm.Get_rank()
        self.c = int(c)
        if self.c < 1 or self.size % self.c:
            raise ValueError("The number of processes must be divisible by the replication factor.")
        self.p_div_c = self.size // self.c
        if self.p_div_c % self.c:
            raise ValueError("The number of processes must be divisible by the square of the replication factor.")
        self.rounds = self.p_div_c // self.c
        self.x, self.y = divmod(self.rank, self.c)

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def Get_coords(self, rank: int):
        return list(divmod(int(rank), self.c))

    def Get_cart_rank(self, coords) -> int:
        return int(coords[0]) * self.c + int(coords[1])

    def Get_topo(self):
        return [self.p_div_c, self.c], [0, 0], [self.x, self.y]

    def Barrier(self):
        self.comm.Barrier()

    def row_ranks(self) -> List[int]:
        """the ``c`` ranks that hold the same rows of ``Y`` (the reference's ``reduce_comm``)"""
        return [self.x * self.c + j for j in range(self.c)]


def largest_power_of_two_square(x: int) -> int:
    """default replication factor: the largest power of two whose square is at most ``x`` (``spmm_15d_main.py:82-90``)"""
    return 2 ** int(math.floor(math.log2(x) / 2))


def _row_block(A, r0: int, r1: int, n_cols: int) -> sparse.csr_matrix:
    """rows ``[r0, r1)`` of a SciPy matrix or of a ``(data, indices, indptr)`` triplet (``:242-256``)"""
    if isinstance(A, tuple):
        data, indices, indptr = A
        ip = np.asarray(indptr[r0:r1 + 1]).astype(np.int64)
        a, b = int(ip[0]), int(ip[-1])
        return sparse.csr_matrix((np.asarray(data[a:b], dtype=np.float32), np.asarray(indices[a:b]), ip - a),
                                 shape=(r1 - r0, n_cols))
    return sparse.csr_matrix(A[r0:r1])


def _shape_of(A) -> Tuple[int, int, int]:
    if isinstance(A, tuple):
        n = int(A[2].shape[0]) - 1
        return n, n, int(A[0].shape[0])
    return int(A.shape[0]), int(A.shape[1]), int(A.nnz)


def _decompose(A, X_cols: int, dtype, c: int, rng, comm, X_full: Optional[np.ndarray]):
    comm = comm if comm is not None else comm_mod.world_comm()
    grid = Grid15D(comm, c)
    if np.dtype(dtype) != np.float32:
        raise ValueError("the device path computes in float32 (the reference's default --type)")
    sizes = comm.bcast(_shape_of(A) if grid.rank == 0 else None, 0)
    NI, NK, _ = sizes
    lNI, lNKb = int(np.ceil(NI / grid.p_div_c)), int(np.ceil(NK / grid.p_div_c))
    lNK = lNKb * grid.rounds

    def block_of(src, x, y):
        rows = _row_block(src, min(NI, x * lNI), min(NI, (x + 1) * lNI), NK)
        b = sparse.csr_matrix(rows[:, min(NK, y * lNK):min(NK, (y + 1) * lNK)], dtype=np.float32)
        b.sum_duplicates()
        b.sort_indices()
        return b

    everyone_has_A = all(comm.allgather(A is not None))
    if everyone_has_A:                         # every rank cuts its own block (memory-mapped files: no root bottleneck)
        lA = block_of(A, grid.x, grid.y)
    else:                                      # the reference's way: rank 0 cuts and sends (:87-118)
        blocks = None
        if grid.rank == 0:
            blocks = [block_of(A, *divmod(r, grid.c)) for r in range(grid.size)]
        lA = comm.bcast(blocks, 0)[grid.rank]
    lA_blocks = []
    for r in range(grid.rounds):
        blk = sparse.csr_matrix(lA[:, min(lA.shape[1], r * lNKb):min(lA.shape[1], (r + 1) * lNKb)])
        blk.sum_duplicates()
        blk.sort_indices()
        lA_blocks.append(blk)
    # X block x is replicated along grid row x; generated by the rank with y == 0 (:134-152)
    actual = max(min(NK, (grid.x + 1) * lNKb) - grid.x * lNKb, 0)
    if X_full is not None:
        X = np.ascontiguousarray(X_full[grid.x * lNKb: grid.x * lNKb + actual], dtype=np.float32)
    else:
        mine = synth.generate_dense_matrix(actual, X_cols, np.float32, rng) if grid.y == 0 else None
        X = comm.allgather(mine)[grid.x * grid.c]
    Y = np.empty((lA.shape[0], X_cols), dtype=np.float32)
    return lA_blocks, X, Y, grid, grid, grid, lNKb


def generate_15d_decomposition(A, X_cols: int, dtype, c: int, rng: np.random.Generator, comm=None,
                               X_full: Optional[np.ndarray] = None):
    """``(lA_blocks, X, Y, cart, bcast, reduce, lNKb)`` like the reference (``:19-154``); the three communicator slots
    hold the same ``Grid15D``.  ``A`` may be given on rank 0 only (the reference) or on every rank (each cuts its own
    block).  ``X_full`` (optional, tests) replaces the random features."""
    return _decompose(A, X_cols, dtype, c, rng, comm, X_full)


def generate_15d_decomposition_new(A, X_cols: int, dtype, c: int, rng: np.random.Generator, comm=None,
                                   X_full: Optional[np.ndarray] = None):
    """Same from a memory-mapped ``(data, indices, indptr)`` triplet (``:157-310``)."""
    return _decompose(A, X_cols, dtype, c, rng, comm, X_full)


class Spmm15D:
    """One rank's resident state of the 1.5D product."""

    def __init__(self, grid: Grid15D, lA_blocks: Sequence[sparse.csr_matrix], x_rows: int, k: int, device: int = 0,
                 fabric=None):
        self.grid, self.k = grid, int(k)
        self.rounds = grid.rounds
        assert len(lA_blocks) == self.rounds
        self.lNI = int(lA_blocks[0].shape[0])
        self.x_rows = int(x_rows)
        self.block_cols = [int(b.shape[1]) for b in lA_blocks]
        self.fab = fabric if fabric is not None else CudaHaloFabric(grid.comm, device)
        tiles = dict(X=self.x_rows, buf=max(self.block_cols + [0]), P=self.lNI)
        if grid.c > 1:
            tiles["Y"] = self.lNI
        self.fab.alloc(tiles, self.k)
        self.A = [self.fab.csr_upload(sparse.csr_matrix(b)) for b in lA_blocks]
        self.nnz = int(sum(b.nnz for b in lA_blocks))
        self.out = "Y" if grid.c > 1 else "P"
        q_own = [grid.y * self.rounds + r for r in range(self.rounds)]
        self.owners = [q * grid.c + grid.y for q in q_own]          # world rank (q, y) holding X block q

    def set_features(self, X: np.ndarray):
        X = np.ascontiguousarray(X, dtype=np.float32)
        if X.shape != (self.x_rows, self.k):
            raise ValueError(f"expected the local X block of shape {(self.x_rows, self.k)}, got {X.shape}")
        if self.x_rows:
            self.fab.h2d("X", 0, X)

    def spmm(self):
        fab, g = self.fab, self.grid
        fab.barrier()                                  # every rank's X tile is final
        for r in range(self.rounds):
            rows, owner = self.block_cols[r], self.owners[r]
            if owner == g.rank:
                src = "X"
            else:
                fab.pull("buf", 0, owner, "X", 0, rows)
                src = "buf"
            fab.spmm(self.A[r], src, "P", accumulate=r > 0)
        if g.c > 1:
            fab.barrier()                              # all partial results are final
            for j, peer in enumerate(g.row_ranks()):
                if j == 0:
                    fab.pull("Y", 0, peer, "P", 0, self.lNI)
                else:
                    fab.accumulate_from("Y", peer, "P", self.lNI)
        fab.barrier()                                  # nobody still reads this rank's X / P tiles

    def result(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if self.lNI == 0:
            self.fab.sync()
            return np.zeros((0, self.k), np.float32) if out is None else out
        return self.fab.d2h(self.out, 0, self.lNI, out)

    def synchronize(self):
        self.fab.sync()

    def flops(self) -> float:
        return 2.0 * self.nnz * self.k

    def close(self):
        if hasattr(self.fab, "close"):
            self.fab.close()


def spmm_15d_cpu(A, X, Y, cart_comm, bcast_comm=None, reduce_comm=None):
    raise NotImplementedError("arrow_matrix_b200 has no CPU arithmetic; use spmm_15d_gpu (the reference's CPU path is "
                              "arrow/baseline/spmm_15d.py:313-367)")


def spmm_15d_gpu(A: List[sparse.csr_matrix], X: np.ndarray, Y: np.ndarray, cart_comm: Grid15D, bcast_comm=None,
                 reduce_comm=None, max_rows: Optional[int] = None, block_size: Optional[int] = None,
                 device: Optional[int] = None) -> np.ndarray:
    """Host-array entry with the reference's signature (``:370-449``): ``Y[:] = (A X)`` rows of this grid row.  Device
    state is created on first use and cached on the grid object; the tiling arguments are accepted and ignored."""
    grid = cart_comm
    k = X.shape[1]
    eng = getattr(grid, "_engine", None)
    if eng is None or eng.k != k:
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        eng = Spmm15D(grid, A, X.shape[0], k, device=device)
        grid._engine = eng
    eng.set_features(X)
    eng.spmm()
    eng.result(Y)
    return Y


def benchmark_15d(A, columns: int, replication: int, iterations: int, device: str, rng: np.random.Generator,
                  validate: bool = True, new_decomposition: bool = False, dataset_name: str = "random", comm=None,
                  wandb_api_key: Optional[str] = None, verbose: bool = True, fabric=None):
    """The reference's driver (``scripts/spmm_15d_main.py:131-281``): decomposition, optional validation against
    ``A @ X`` on rank 0, ``iterations`` timed products.  Returns ``dict(engine, times, validation)``."""
    comm = comm if comm is not None else comm_mod.world_comm()
    rank, size = comm.Get_rank(), comm.Get_size()
    c = replication if replication else largest_power_of_two_square(size)
    func = generate_15d_decomposition_new if new_decomposition else generate_15d_decomposition
    lA, lX, lY, grid, _, _, lNKb = func(A, columns, np.float32, c, rng, comm=comm)
    if device != "gpu":
        raise NotImplementedError("arrow_matrix_b200 has no CPU arithmetic; use device='gpu'")
    eng = Spmm15D(grid, lA, lX.shape[0], columns, device=int(os.environ.get("LOCAL_RANK", "0")), fabric=fabric)
    grid._engine = eng
    validation = None
    if validate:
        spmm_15d_gpu(lA, lX, lY, grid)
        xs = comm.allgather(lX if grid.y == 0 else None)
        ys = comm.allgather(lY if grid.y == 0 else None)
        if rank == 0:
            Afull = sparse.csr_matrix(A, dtype=np.float32) if not isinstance(A, tuple) else \
                sparse.csr_matrix((np.asarray(A[0], dtype=np.float32), np.asarray(A[1]), np.asarray(A[2])),
                                  shape=(A[2].shape[0] - 1,) * 2)
            X = np.concatenate([x for x in xs if x is not None])[: Afull.shape[1]]
            Yg = np.concatenate([y for y in ys if y is not None])[: Afull.shape[0]]
            ref = Afull @ X
            validation = float(np.linalg.norm(Yg - ref) / max(np.linalg.norm(ref), 1e-30))
            if verbose:
                print(f"GPU validation: {np.allclose(Yg, ref, rtol=1e-4, atol=1e-5)} ({validation})", flush=True)
    wb_logging.wandb_init(comm, dataset_name, columns, iterations, "gpu", f"15D_B200_c_{c}_v0.1", lX.shape[0], wandb_api_key)
    wb_logging.set_iteration_data({"gpu_tiling": False})
    times = []
    for i in range(iterations):
        eng.set_features(lX)
        eng.synchronize()
        comm.Barrier()
        tic = time.perf_counter()
        eng.spmm()
        eng.synchronize()
        toc = time.perf_counter()
        times.append(toc - tic)
        wb_logging.log({"spmm_time": toc - tic, "iteration": i})
    if verbose and rank == 0 and times:
        print(f"GPU: {1e3 * float(np.median(times)):.3f} ms +- {1e3 * float(np.std(times)):.3f}", flush=True)
    wb_logging.finish(comm)
    comm.Barrier()
    return dict(engine=eng, times=times, validation=validation)
