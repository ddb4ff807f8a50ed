"""ORACLE / CPU BASELINE -- test and measurement infrastructure only (never used by the product).

The reference's CPU path of one iteration, re-hosted on the cores of one machine.  The literal
reference cannot be launched here (no mpi4py / mpiexec, and it needs one MPI rank per block-row:
>= 1000 ranks for the 10M-row workload, arrow/arrow_bench.py:70-78), so its arithmetic and data
movement are reproduced with the same split it uses -- independent groups of block-rows -- on
``P`` host threads:

  * products: the restated SciPy ``csr_matvecs`` kernel (oracle/csr_matvecs.c, bit-identical to
    ``scipy csr @ dense`` = arrow_slim_mpi.py:109-111, 125-127, 142-144) on row ranges; ctypes
    releases the GIL, so threads run in parallel the way MPI ranks would;
  * forward / backward exchange: row gather and scatter-add (arrow_dec_mpi.py:526/544 and 421/437).

Used by bench.py for the ``cpu_baseline`` object and the ``--impl reference`` arm.
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import List, Sequence, Tuple

import numpy as np
from scipy import sparse

from . import oracle as _o


def _nnz_balanced_ranges(indptr: np.ndarray, parts: int) -> List[Tuple[int, int]]:
    n = indptr.size - 1
    total = int(indptr[-1] - indptr[0])
    if n == 0:
        return []
    targets = indptr[0] + (np.arange(1, parts, dtype=np.int64) * total) // parts
    cuts = np.searchsorted(indptr, targets, side="left")
    bounds = np.unique(np.concatenate([[0], cuts, [n]])).astype(np.int64)
    return [(int(bounds[i]), int(bounds[i + 1])) for i in range(bounds.size - 1) if bounds[i + 1] > bounds[i]]


class CpuArrowReference:
    def __init__(self, decomposition: Sequence[Tuple[sparse.csr_matrix, np.ndarray]], width: int, k: int,
                 n_threads: int = 0, tasks_per_thread: int = 4):
        self.lib = _o._lib()
        self.k, self.width = k, width
        self.P = n_threads or (os.cpu_count() or 1)
        self._cpus = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(self.P))
        # persistent workers with a STATIC task -> thread map (task i runs on thread i mod P), one thread per core:
        # the same thread first-touches and later computes a row range, like an MPI rank owning its block-rows
        import threading
        self._start = threading.Barrier(self.P + 1)
        self._done = threading.Barrier(self.P + 1)
        self._job = None
        self._stop = False
        self._errors = []
        self._threads = [threading.Thread(target=self._worker, args=(t,), daemon=True) for t in range(self.P)]
        for th in self._threads:
            th.start()
        self.L = len(decomposition)
        self.n_blocks = [_o.number_of_blocks(B, width) for B, _ in decomposition]
        _, self.to_prev, _, _ = _o.prepare_permutations([p for _, p in decomposition], self.n_blocks, width)
        self.rows = [b * width for b in self.n_blocks]
        self.levels = []
        for (B, _), rows in zip(decomposition, self.rows):
            B = sparse.csr_matrix(B)
            nnz = int(B.indptr[rows])
            idx = np.ascontiguousarray(B.indices[:nnz], dtype=np.int32)
            if nnz and int(idx.max()) >= rows:
                raise ValueError("CPU baseline expects arrow-shaped levels (columns inside the active block-rows)")
            self.levels.append(dict(indptr=np.ascontiguousarray(B.indptr[:rows + 1], dtype=np.int32), indices=idx,
                                    data=np.ascontiguousarray(B.data[:nnz], dtype=np.float32), nnz=nnz))
        self.maps = [None] + [np.ascontiguousarray(self.to_prev[j][: self.rows[j]], dtype=np.int64) for j in range(1, self.L)]
        nt = self.P * tasks_per_thread
        self.spmm_ranges = [_nnz_balanced_ranges(lv["indptr"], nt) for lv in self.levels]
        self.row_ranges = [[(int(a), int(b)) for a, b in zip(np.linspace(0, r, nt + 1)[:-1].astype(np.int64),
                                                             np.linspace(0, r, nt + 1)[1:].astype(np.int64)) if b > a]
                           for r in self.rows]
        self.total_nnz = sum(lv["nnz"] for lv in self.levels)
        self.X = [np.empty((r, k), np.float32) for r in self.rows]
        self.C = [np.empty((r, k), np.float32) for r in self.rows]
        for j in range(self.L):
            self._first_touch(self.X[j], self.spmm_ranges[j])
            self._first_touch(self.C[j], self.spmm_ranges[j])

    def _worker(self, t: int):
        try:
            os.sched_setaffinity(0, {self._cpus[t % len(self._cpus)]})      # acts on the calling thread
        except Exception:
            pass
        while True:
            self._start.wait()
            if self._stop:
                return
            fn, tasks = self._job
            try:
                for i in range(t, len(tasks), self.P):
                    fn(tasks[i])
            except BaseException as e:      # noqa: BLE001
                self._errors.append(e)
            self._done.wait()

    def _run(self, fn, tasks):
        self._job = (fn, tasks)
        self._start.wait()
        self._done.wait()
        if self._errors:
            raise self._errors.pop()

    def _first_touch(self, arr, ranges):
        """pages are placed on the NUMA node of the thread that first writes them: let each worker zero the row
        ranges it will later compute, the way every MPI rank of the reference allocates its own tiles"""
        k = arr.shape[1]
        lib, P_, i64 = self.lib, ctypes.c_void_p, ctypes.c_int64
        self._run(lambda rg: lib.oracle_zero_rows_f32(i64(rg[0]), i64(rg[1]), i64(k), P_(arr.ctypes.data)), ranges)

    def set_features(self, X0: np.ndarray):
        """copy into the NUMA-placed level-0 tile (parallel, by row ranges)"""
        X0 = np.ascontiguousarray(X0, dtype=np.float32)
        dst = self.X[0]

        def cp(rg):
            dst[rg[0]:rg[1]] = X0[rg[0]:rg[1]]
        self._run(cp, self.spmm_ranges[0])

    def step(self) -> np.ndarray:
        lib, k, P_ = self.lib, self.k, ctypes.c_void_p
        i64 = ctypes.c_int64
        # forward exchange: X_j[r] = X_{j-1}[to_prev_j[r]]
        for j in range(1, self.L):
            src, dst, m = self.X[j - 1], self.X[j], self.maps[j]
            self._run(lambda rg: lib.oracle_gather_rows_f32_range(i64(rg[0]), i64(rg[1]), i64(k), P_(m.ctypes.data),
                                                                  i64(src.shape[0]), P_(src.ctypes.data), P_(dst.ctypes.data)),
                      self.row_ranges[j])
        # products per level (zero + accumulate like scipy's `zeros(); csr_matvecs()`)
        for j in range(self.L):
            lv, X, C = self.levels[j], self.X[j], self.C[j]

            def task(rg, lv=lv, X=X, C=C):
                lib.oracle_zero_rows_f32(i64(rg[0]), i64(rg[1]), i64(k), P_(C.ctypes.data))
                lib.oracle_csr_matvecs_f32_i32_rows(i64(rg[0]), i64(rg[1]), i64(k), P_(lv["indptr"].ctypes.data),
                                                    P_(lv["indices"].ctypes.data), P_(lv["data"].ctypes.data),
                                                    P_(X.ctypes.data), P_(C.ctypes.data))
            self._run(task, self.spmm_ranges[j])
        # backward exchange: C_{j-1}[to_prev_j[r]] += C_j[r]
        for j in range(self.L - 1, 0, -1):
            src, dst, m = self.C[j], self.C[j - 1], self.maps[j]
            self._run(lambda rg: lib.oracle_scatter_add_rows_f32_range(i64(rg[0]), i64(rg[1]), i64(k), P_(m.ctypes.data),
                                                                       i64(dst.shape[0]), P_(src.ctypes.data), P_(dst.ctypes.data)),
                      self.row_ranges[j])
        return self.C[0]

    def time_steps(self, steps: int, warmup: int = 1) -> float:
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return (time.perf_counter() - t0) / max(steps, 1)

    def flops_per_step(self) -> float:
        return 2.0 * self.total_nnz * self.k

    def close(self):
        if not self._stop:
            self._stop = True
            self._start.wait()
            for th in self._threads:
                th.join(5)
