"""PETSc-style 1D row-partitioned SpMM with a halo exchange, on the library's kernels (SURVEY.md "next" row N4).

Mirror of the reference's ``arrow/baseline/spmm_petsc.py``: every rank (here: every GPU) owns ``n_i`` consecutive
rows of ``A`` and of ``X``; ``Y_i = A_i_local X_i + A_i_nonlocal X_nonlocal`` where ``X_nonlocal`` are the rows of
other ranks' ``X`` that ``A_i`` references (tables: ``arrow_matrix_b200.matrix_slice.MatrixSlice``).

What changes against the reference's GPU path (``:229-326``): the two blocks are uploaded once instead of per call
(``_sp2cp`` at ``:258, 303``), ``X`` stays on the device, and the exchange is a device-side pack
(``arrow_gather_rows``: rows grouped by destination, in the destination's halo order, so the receiver's part is
contiguous) followed by one peer copy per source over NVLink -- instead of ``Isend/Irecv`` of host arrays
(``:112-144``).  The exchange runs on the side lane while the local product runs on the main lane, like the
reference overlaps its non-blocking messages with ``A_i_local @ X_i`` (``:199-214``).

Reference-facing functions keep their names and arguments (``spmm_gpu``, ``load_matrix_slice``,
``benchmark_spmm``); ``spmm_cpu`` raises -- this package has no CPU arithmetic.
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional

import numpy as np
from scipy import sparse

from .. import comm as comm_mod
from .. import synth, wb_logging
from ..matrix_slice import MatrixSlice


class CudaHaloFabric:
    """Device tiles of one rank in a single IPC-exported arena + the peers' views of it.

    Arena layout (floats): ``[0,64)`` barrier flags of the main lane, ``[64,128)`` flags of the side lane, then the
    tiles in the order given to ``alloc`` (each aligned to 64 floats)."""

    SIDE = 3

    def __init__(self, comm, device: int = 0, stream: Optional[int] = None):
        from .. import _lib
        self._lib = _lib
        self.comm = comm
        self.rank, self.world = comm.Get_rank(), comm.Get_size()
        if self.world > 16:
            raise ValueError("the device-side peer barrier supports at most 16 GPUs")
        self.ctx = _lib.Context(device, stream)
        self._side = False
        self._ident = {}

    def alloc(self, rows: Dict[str, int], k: int):
        ctx = self.ctx
        self.k = int(k)
        ctx.preload_kernels(self.k)        # nothing may be loaded for the first time while a peer barrier spins
        offs, pos = {}, 128
        for name, r in rows.items():
            offs[name] = pos
            pos += -(-(int(r) * self.k) // 64) * 64
        arena_rows = max(-(-pos // 64), (2 << 20) // 256)          # >= 2 MiB: the driver gives it a block of its own
        self._arena = ctx.dense_alloc(arena_rows, 64)
        ctx.sync()
        mine = dict(handle=self._arena.ipc_export(), arena_rows=arena_rows, offs=offs, rows={n: int(r) for n, r in rows.items()})
        everyone = self.comm.allgather(mine)
        self._info = everyone
        self._base, self._flags, self._flags_side, self._keep = [], [], [], []
        for g, info in enumerate(everyone):
            arena = self._arena if g == self.rank else ctx.ipc_import(info["handle"], info["arena_rows"], 64)
            self._keep.append(arena)
            base = arena.device_ptr()
            self._base.append(base)
            self._flags.append(ctx.dense_wrap(base, 1, 64))
            self._flags_side.append(ctx.dense_wrap(base + 64 * 4, 1, 64))
        self._tiles = {n: ctx.dense_wrap(self._base[self.rank] + offs[n] * 4, int(r), self.k) for n, r in rows.items()}
        self._views = {}

    def tile(self, name: str):
        return self._tiles[name]

    def view(self, g: int, name: str, row0: int, rows: int):
        """rows ``[row0, row0+rows)`` of tile ``name`` of rank ``g`` (peer memory when ``g`` is not this rank)"""
        key = (g, name, row0, rows)
        v = self._views.get(key)
        if v is None:
            v = self.ctx.dense_wrap(self._base[g] + (self._info[g]["offs"][name] + row0 * self.k) * 4, rows, self.k)
            self._views[key] = v
        return v

    def csr_upload(self, A: sparse.csr_matrix):
        A = sparse.csr_matrix(A)
        return self.ctx.csr_upload(A.shape[0], A.shape[1], A.indptr, A.indices, A.data.astype(np.float32, copy=False))

    def map_upload(self, m, limit):
        return self.ctx.map_upload(np.asarray(m, dtype=np.int64), max(int(limit), 1))

    def h2d(self, name, row0, X):
        self._tiles[name].h2d(X, row0=row0)

    def d2h(self, name, row0, rows, out=None):
        return self._tiles[name].d2h(out, row0=row0, rows=rows)

    def fill(self, name, v):
        self._tiles[name].fill(v)

    def spmm(self, A, x_name, y_name, accumulate=False):
        self.ctx.spmm(A, self._tiles[x_name], self._tiles[y_name], accumulate=accumulate)

    # -- everything below is issued on the side lane between side_begin() and side_join() ----------------
    def side_begin(self):
        self.ctx.lane_wait(self.SIDE, 0)
        self.ctx.set_lane(self.SIDE)
        self._side = True

    def side_join(self):
        self.ctx.set_lane(0)
        self._side = False
        self.ctx.lane_wait(0, self.SIDE)

    def pack(self, dst_name, src_name, row_map):
        if row_map.n:
            self.ctx.gather_rows(self.view(self.rank, dst_name, 0, row_map.n), self._tiles[src_name], row_map)

    def barrier(self):
        if self.world > 1:
            self.ctx.peer_barrier(self._flags_side if self._side else self._flags, self.rank)

    def pull(self, dst_name, dst_row0, peer, src_name, src_row0, rows):
        if rows:
            self.view(self.rank, dst_name, dst_row0, rows).copy_from(self.view(peer, src_name, src_row0, rows), rows=rows)

    def accumulate_from(self, dst_name, peer, src_name, rows):
        """``dst[:rows] += (rank peer's src)[:rows]`` (peer read over NVLink, local read-modify-write)"""
        if rows:
            ident = self._ident.get(rows)
            if ident is None:
                ident = self.ctx.map_upload(np.arange(rows, dtype=np.int64), rows)
                self._ident[rows] = ident
            self.ctx.gather_rows(self.view(self.rank, dst_name, 0, rows), self.view(peer, src_name, 0, rows), ident,
                                 accumulate=True)

    def sync(self):
        if self.world > 1:
            self.ctx.lane_sync(self.SIDE)
        self.ctx.sync()

    def close(self):
        self.ctx.close()


class HaloSpmm:
    """One rank's resident state of the 1D halo-exchange SpMM: ``Y_i = A_i X`` for the rows of a ``MatrixSlice``."""

    def __init__(self, comm, matrix_slice: MatrixSlice, k: int, device: int = 0, fabric=None, overlap: bool = True):
        sl = matrix_slice
        self.comm, self.sl, self.k = comm, sl, int(k)
        self.rank, self.world = comm.Get_rank(), comm.Get_size()
        self.n_i = int(sl.A_i_local.shape[0])
        self.n_halo = int(sl.x_index_in.size)
        self.n_send = int(sl.x_index_out.size)
        self.overlap = bool(overlap) and self.world > 1
        self.fab = fabric if fabric is not None else CudaHaloFabric(comm, device)
        self.fab.alloc(dict(X=self.n_i + self.n_halo, Y=self.n_i, send=self.n_send), self.k)
        n_x = self.n_i + self.n_halo
        non = sparse.csr_matrix(sl.A_i_nonlocal)
        shifted = sparse.csr_matrix((non.data, non.indices.astype(np.int64) + self.n_i, non.indptr), shape=(self.n_i, n_x))
        if self.overlap:
            self.A_loc = self.fab.csr_upload(sparse.csr_matrix(sl.A_i_local))
            self.A_non = self.fab.csr_upload(shifted) if shifted.nnz else None
            self.A_cat = None
        else:
            loc = sparse.csr_matrix(sl.A_i_local)
            loc.resize((self.n_i, n_x))
            cat = (loc + shifted).tocsr() if shifted.nnz else loc
            cat.sum_duplicates()
            cat.sort_indices()
            self.A_cat = self.fab.csr_upload(cat)
        self.nnz = int(sl.A_i_local.nnz + sl.A_i_nonlocal.nnz)
        # pack order: x_index_out is sorted by (destination rank, global row) == the destination's halo order
        self.pack_map = self.fab.map_upload(sl.x_index_out_localized, self.n_i)
        counts = comm.allgather([int(c) for c in sl.send_count])              # counts[s][d]: rows s sends to d
        # where, inside peer s's send tile, the rows meant for this rank start
        self.region = [int(sum(counts[s][: self.rank])) for s in range(self.world)]
        for s in range(self.world):
            assert counts[s][self.rank] == int(sl.recv_count[s]), "send/receive tables disagree"

    def set_features(self, X_i: np.ndarray):
        X_i = np.ascontiguousarray(X_i, dtype=np.float32)
        if X_i.shape != (self.n_i, self.k):
            raise ValueError(f"expected X_i of shape {(self.n_i, self.k)}, got {X_i.shape}")
        if self.n_i:
            self.fab.h2d("X", 0, X_i)

    def _exchange(self):
        """pack -> barrier -> every rank copies its part of each peer's send tile -> barrier (send tiles reusable)"""
        fab, sl = self.fab, self.sl
        fab.pack("send", "X", self.pack_map)
        fab.barrier()
        for g in range(self.world):
            c = int(sl.recv_count[g])
            if c and g != self.rank:
                fab.pull("X", self.n_i + int(sl.recv_sdispl[g]), g, "send", self.region[g], c)
        fab.barrier()

    def spmm(self):
        """``Y_i = A_i X`` (stream-ordered; ``result()`` or ``synchronize()`` waits)."""
        fab = self.fab
        if self.world > 1:
            fab.side_begin()
            self._exchange()
        if self.overlap:
            fab.spmm(self.A_loc, "X", "Y")      # main lane, while the side lane exchanges
            fab.side_join()
            if self.A_non is not None:
                fab.spmm(self.A_non, "X", "Y", accumulate=True)
        else:
            if self.world > 1:
                fab.side_join()
            fab.spmm(self.A_cat, "X", "Y")

    def result(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if self.n_i == 0:
            self.fab.sync()
            return np.zeros((0, self.k), np.float32) if out is None else out
        return self.fab.d2h("Y", 0, self.n_i, out)

    def halo(self) -> np.ndarray:
        """the received rows (the reference's ``X_i_nonlocal`` after the exchange)"""
        if self.n_halo == 0:
            return np.zeros((0, self.k), np.float32)
        return self.fab.d2h("X", self.n_i, self.n_halo)

    def synchronize(self):
        self.fab.sync()

    def flops(self) -> float:
        return 2.0 * self.nnz * self.k

    def close(self):
        if hasattr(self.fab, "close"):
            self.fab.close()


# ---------------------------------------------------------------------------------------------------------
# reference-facing functions
# ---------------------------------------------------------------------------------------------------------
def load_matrix_slice(some_slice: str, rank: int) -> sparse.csr_matrix:
    """``{name}.part.{P}.slice.{y}.npz`` -> the file of this rank (``y`` replaced, ``spmm_petsc.py:84-103``)."""
    parts = some_slice.split(".")
    parts[-2] = str(rank)
    return sparse.load_npz(".".join(parts))


def spmm_cpu(comm, matrix_slice, X_i_local, Y_i_local, X_i_nonlocal):
    raise NotImplementedError("arrow_matrix_b200 has no CPU arithmetic; use device='gpu' (the reference's CPU path is "
                              "arrow/baseline/spmm_petsc.py:183-226)")


def spmm_gpu(comm, matrix_slice: MatrixSlice, X_i_local: np.ndarray, Y_i_local: np.ndarray, X_i_nonlocal: np.ndarray,
             bsize_local: Optional[int] = None, bsize_nonlocal: Optional[int] = None, device: int = 0) -> np.ndarray:
    """Host-array entry with the reference's signature and effects (``:229-326``): ``Y_i_local += A_i X`` and
    ``X_i_nonlocal`` receives the halo rows.  The device state is created on first use and cached on the slice; the
    column-tiling arguments are accepted and ignored (180 GB of HBM: no tiling needed at the reference's sizes)."""
    k = X_i_local.shape[1]
    eng = getattr(matrix_slice, "_halo_engine", None)
    if eng is None or eng.k != k:
        eng = HaloSpmm(comm, matrix_slice, k, device=device)
        matrix_slice._halo_engine = eng
    eng.set_features(X_i_local)
    eng.spmm()
    Y_i_local += eng.result()
    if X_i_nonlocal is not None and X_i_nonlocal.size:
        X_i_nonlocal[:] = eng.halo()
    return Y_i_local


def benchmark_spmm(matrix_slice_file: Optional[str], k: int, iterations: int, device: str, wandb_api_key: Optional[str],
                   dtype, rng: np.random.Generator, gpu_tiling: bool = False, dryrun: bool = False,
                   mem_fraction: float = 0.9, comm=None, scale: int = 4 * 1024, verbose: bool = True,
                   device_id: Optional[int] = None):
    """The reference's driver (``:389-495``): slice per rank (file or synthetic), tables, ``iterations`` products on
    fresh features, ``spmm_time`` logged per iteration.  Returns ``dict(engine, times)`` (the reference returns None)."""
    comm = comm if comm is not None else comm_mod.world_comm()
    rank, size = comm.Get_rank(), comm.Get_size()
    if np.dtype(dtype) != np.float32:
        raise ValueError("the device path computes in float32 (the reference's default --type)")
    dataset = matrix_slice_file.split(".")[0] if matrix_slice_file is not None else None
    wb_logging.wandb_init(comm, dataset, k, iterations, device, "PETSc_B200_v0.1", 0, wandb_api_key)
    if matrix_slice_file is None:
        A_i = synth.generate_sparse_matrix(scale, size * scale, scale * 10, np.float32, rng)      # :416-418
    else:
        nr_parts = int(matrix_slice_file.split(".")[-4])
        if nr_parts != size:
            raise ValueError(f"Number of parts in file name ({nr_parts}) does not match number of ranks ({size})")
        A_i = sparse.csr_matrix(load_matrix_slice(matrix_slice_file, rank).astype(np.float32))
        A_i.eliminate_zeros()
        A_i.sort_indices()
        A_i.sum_duplicates()
    mat_slice = MatrixSlice.initialize(comm, A_i)
    wb_logging.log({"nonlocal_columns": mat_slice.A_i_nonlocal.shape[1], "local_columns": mat_slice.A_i_local.shape[1]})
    if dryrun:
        return None
    if device != "gpu":
        raise NotImplementedError("arrow_matrix_b200 has no CPU arithmetic; use device='gpu'")
    if device_id is None:                      # one process per GPU: torchrun exports LOCAL_RANK
        device_id = int(os.environ.get("LOCAL_RANK", "0"))
    eng = HaloSpmm(comm, mat_slice, k, device=device_id)
    wb_logging.set_iteration_data({"gpu_tiling": False})
    times: List[float] = []
    for i in range(iterations):
        wb_logging.set_iteration_data({"iteration": i})
        X_i = synth.generate_dense_matrix(A_i.shape[0], k, np.float32, rng)
        eng.set_features(X_i)
        eng.synchronize()
        comm.Barrier()
        tic = time.perf_counter()
        ok = True
        try:
            eng.spmm()
            eng.synchronize()
        except Exception as e:     # noqa: BLE001 - collective abort like the reference (:479-489)
            print(f"Rank {rank} encountered an error: {e}", flush=True)
            ok = False
        toc = time.perf_counter()
        wb_logging.log({"spmm_time": toc - tic})
        times.append(toc - tic)
        if verbose:
            print("RANK", rank, "Iteration", i, " -- ", toc - tic, "s", flush=True)
        if comm.allreduce_lor(not ok):
            break
    wb_logging.finish(comm)
    comm.Barrier()
    return dict(engine=eng, times=times)
