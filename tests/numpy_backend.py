"""Test double for the sharded engine's backend: numpy/scipy tiles, gloo for the peers (CPU only).

Peer memory is emulated faithfully to the protocol's own rule -- a rank may only read peer data that
was final before the last barrier: at every ``barrier()`` each rank publishes a snapshot of its tiles
(all_gather over gloo) and remote reads are served from the snapshots.  A missing barrier in the engine
therefore shows up as a stale read and a failing test.
"""
import numpy as np
from scipy import sparse


class _Map:
    def __init__(self, m, limit):
        m = np.asarray(m, dtype=np.int64)
        self.m = np.where((m < 0) | (m >= limit), -1, m)
        self.n = m.size


class GlooNumpyBackend:
    def __init__(self, comm, width, plan=None):
        self.comm, self.width, self.plan = comm, width, plan
        self.rank, self.world = comm.Get_rank(), comm.Get_size()
        self.ctx = None
        self.snap = None
        self.n_barriers = 0

    def csr_upload(self, n_rows, n_cols, indptr, indices, data):
        return sparse.csr_matrix((data, indices, indptr), shape=(n_rows, n_cols), dtype=np.float32)

    def map_upload(self, m, limit):
        return _Map(m, limit)

    def alloc_shared_tiles(self, rows_per_level, k, tiles_per_level=None):
        self.k = k
        tiles_per_level = tiles_per_level or [2] * len(rows_per_level)
        self.tiles = [[np.zeros((r, k), np.float32) for _ in range(nt)] for r, nt in zip(rows_per_level, tiles_per_level)]
        return self.tiles

    def h2d(self, tile, off, X):
        tile[off:off + X.shape[0]] = X

    def d2h(self, tile, off, rows, out=None):
        if out is None:
            return tile[off:off + rows].copy()
        out[:] = tile[off:off + rows]
        return out

    def sync(self):
        pass

    def fill(self, tile, v):
        tile[:] = v

    def side_begin(self):
        pass

    def limit_spmm(self, n):
        pass

    def side_join(self):
        pass

    def barrier(self, side=False):
        # remote stores issued before the barrier are visible after it: deliver the outboxes first ...
        outbox = getattr(self, "outbox", [])
        self.outbox = []
        for sender in self.comm.allgather(outbox):
            for dest, level, which, rows, data in sender:
                if dest == self.rank:
                    self.tiles[level][which][rows] = data
        # ... then publish what peers may read until the next barrier
        send = getattr(self, "sendbuf", None)
        got = self.comm.allgather(([[t.copy() for t in pair] for pair in self.tiles], None if send is None else send.copy()))
        self.snap = [g[0] for g in got]
        self.snap_send = [g[1] for g in got]
        self.n_barriers += 1

    # -- fused step primitives, mirroring CudaPeerBackend (remote stores = outbox entries delivered at the next barrier) --
    supports_fused = True

    def fused_matrix(self, A, colmap, n_cols):
        A = sparse.csr_matrix(A)
        idx = np.asarray(colmap, dtype=np.int64)[A.indices]
        assert np.all(idx >= 0) and np.all(idx < n_cols)
        return sparse.csr_matrix((A.data, idx, A.indptr), shape=(A.shape[0], n_cols), dtype=np.float32)

    def out_table(self, tiles, which, row):
        return dict(tiles=list(tiles), which=np.asarray(which), row=np.asarray(row, dtype=np.int64))

    def _store(self, table, values):
        """rows of ``values`` to wherever the table routes them"""
        which, row = table["which"], table["row"]
        for t, (g, lv, ix) in enumerate(table["tiles"]):
            sel = np.flatnonzero(which[: values.shape[0]] == t)
            if sel.size == 0:
                continue
            if g == self.rank:
                self.tiles[lv][ix][row[sel]] = values[sel]
            else:
                self.__dict__.setdefault("outbox", []).append((g, lv, ix, row[sel].copy(), values[sel].copy()))

    def push_plan(self, recv, src_rows, bounds, offs, dests, src_limit):
        return dict(recv=recv, src=np.asarray(src_rows, dtype=np.int64), bounds=[int(b) for b in bounds], offs=[int(o) for o in offs],
                    dests=[int(d) for d in dests])

    def push(self, pp, x, side=False):
        X = self.tiles[x[0]][x[1]]
        for i, d in enumerate(pp["dests"]):
            a, b = pp["bounds"][i], pp["bounds"][i + 1]
            assert d != self.rank and b > a
            rows = pp["offs"][i] + np.arange(b - a, dtype=np.int64)
            self.__dict__.setdefault("outbox", []).append((d, pp["recv"][0], pp["recv"][1], rows, X[pp["src"][a:b]].copy()))

    def spmm_fused(self, A, x, recv, x_split, table, add=None, add_map=None, side=False):
        Xc = np.concatenate([self.tiles[x[0]][x[1]][:x_split], self.tiles[recv[0]][recv[1]]])
        prod = (A @ Xc[: A.shape[1]]).astype(np.float32)
        if add is not None:
            m = add_map.m[: prod.shape[0]]
            sel = np.flatnonzero(m >= 0)
            prod[sel] += self.tiles[add[0]][add[1]][m[sel]]
        self._store(table, prod)

    def reduce_rows(self, tile, rows, table=None, side=False):
        total = np.zeros((rows, self.k), np.float32)
        for g in range(self.world):
            total += self._tile(g, tile[0], tile[1])[:rows]
        if table is None:
            self.tiles[tile[0]][tile[1]][:rows] = total
        else:
            self._store(table, total)

    def final_add(self, dst, dst_off, rows, stage, add_map):
        m = add_map.m[:rows]
        sel = np.flatnonzero(m >= 0)
        self.tiles[dst[0]][dst[1]][dst_off + sel] += self.tiles[stage[0]][stage[1]][m[sel]]

    # -- staged (two-phase) packed exchange, mirroring CudaPeerBackend.stage_rows / apply_staged -----------------
    supports_staged_exchange = True

    def tile_view(self, level, which, off, rows):
        return self.tiles[level][which][off:off + rows]

    def _xtable(self, dst_level, forward):
        key = (dst_level, forward)
        cache = self.__dict__.setdefault("_xt", {})
        if key not in cache:
            raw = self.plan.a2a_tables(dst_level, forward)
            counts = self.comm.allgather([int(c) for c in raw["send_counts"]])
            cache[key] = dict(raw=raw, region=[sum(counts[s][:self.rank]) for s in range(self.world)])
        return cache[key]

    def stage_rows(self, dst_level, src, forward, stage, side=False):
        t = self._xtable(dst_level, forward)
        raw = t["raw"]
        sh = self.plan.levels[src[0]]
        self.sendbuf = self.tiles[src[0]][src[1]][sh.hoff + raw["pack"]].copy()          # pack (local)
        self.barrier(side)                                                                # every peer's pack is complete
        parts = [self.snap_send[g][t["region"][g]: t["region"][g] + int(raw["recv_counts"][g])] if g != self.rank
                 else self.sendbuf[t["region"][g]: t["region"][g] + int(raw["recv_counts"][g])] for g in range(self.world)]
        staged = np.concatenate(parts) if parts else np.zeros((0, self.k), np.float32)
        self.tiles[stage[0]][stage[1]][: staged.shape[0]] = staged

    def apply_staged(self, dst, dst_off, dst_level, forward, stage, accumulate):
        unpack = self._xtable(dst_level, forward)["raw"]["unpack"]
        sel = np.flatnonzero(unpack >= 0)
        rows = self.tiles[stage[0]][stage[1]][unpack[sel]]
        d = self.tiles[dst[0]][dst[1]]
        if accumulate:
            d[dst_off + sel] += rows
        else:
            d[dst_off + sel] = rows

    def allreduce_sum(self, v):
        return sum(self.comm.allgather(int(v)))

    def spmm(self, A, X, C):
        C[:] = A @ X[: A.shape[1]]

    def _tile(self, g, level, which):
        return self.tiles[level][which] if g == self.rank else self.snap[g][level][which]

    def pull_rows(self, dst, dst_off, src, src_bounds, row_map, accumulate, forward=True, side=False):
        d = self.tiles[dst[0]][dst[1]]
        m = row_map.m
        for g in range(self.world):
            sel = np.flatnonzero((m >= src_bounds[g]) & (m < src_bounds[g + 1]))
            if sel.size == 0:
                continue
            hoff = self.plan.hoff_of(src[0], g) if self.plan is not None else (self.width if g > 0 else 0)
            rows = self._tile(g, src[0], src[1])[hoff + m[sel] - src_bounds[g]]
            if accumulate:
                d[dst_off + sel] += rows
            else:
                d[dst_off + sel] = rows

    def copy_rows_from_peer(self, dst, dst_off, peer, src, src_off, rows, side=False):
        self.tiles[dst[0]][dst[1]][dst_off:dst_off + rows] = self._tile(peer, src[0], src[1])[src_off:src_off + rows]

    def bcast_head(self, tile, rows):
        if self.rank > 0:
            self.tiles[tile[0]][tile[1]][:rows] = self._tile(0, tile[0], tile[1])[:rows]

    def reduce_head(self, tile, rows):
        if self.rank == 0:
            for g in range(1, self.world):
                self.tiles[tile[0]][tile[1]][:rows] += self._tile(g, tile[0], tile[1])[:rows]


class GlooNumpyHaloFabric:
    """Test double of ``baseline.spmm_petsc.CudaHaloFabric`` (numpy tiles, gloo peers, snapshot-at-barrier reads)."""

    def __init__(self, comm):
        self.comm = comm
        self.rank, self.world = comm.Get_rank(), comm.Get_size()
        self.snap = None
        self.n_barriers = 0
        self.side = False
        self.log = []

    def alloc(self, rows, k):
        self.k = k
        self.tiles = {n: np.zeros((int(r), k), np.float32) for n, r in rows.items()}

    def csr_upload(self, A):
        return sparse.csr_matrix(A, dtype=np.float32)

    def map_upload(self, m, limit):
        return _Map(m, max(int(limit), 1))

    def h2d(self, name, row0, X):
        self.tiles[name][row0:row0 + X.shape[0]] = X

    def d2h(self, name, row0, rows, out=None):
        if out is None:
            return self.tiles[name][row0:row0 + rows].copy()
        out[:] = self.tiles[name][row0:row0 + rows]
        return out

    def fill(self, name, v):
        self.tiles[name][:] = v

    def spmm(self, A, x_name, y_name, accumulate=False):
        self.log.append(("spmm", self.side))
        prod = A @ self.tiles[x_name][: A.shape[1]]
        if accumulate:
            self.tiles[y_name][:] += prod
        else:
            self.tiles[y_name][:] = prod

    def side_begin(self):
        self.side = True

    def side_join(self):
        self.side = False

    def pack(self, dst_name, src_name, row_map):
        assert self.side or self.world == 1
        if row_map.n:
            assert np.all(row_map.m >= 0)
            self.tiles[dst_name][: row_map.n] = self.tiles[src_name][row_map.m]

    def barrier(self):
        if self.world > 1:
            self.snap = self.comm.allgather({n: t.copy() for n, t in self.tiles.items()})
        self.n_barriers += 1

    def _peer_tile(self, peer, name):
        return self.tiles[name] if peer == self.rank else self.snap[peer][name]

    def pull(self, dst_name, dst_row0, peer, src_name, src_row0, rows):
        if rows:
            self.tiles[dst_name][dst_row0:dst_row0 + rows] = self._peer_tile(peer, src_name)[src_row0:src_row0 + rows]

    def accumulate_from(self, dst_name, peer, src_name, rows):
        if rows:
            self.tiles[dst_name][:rows] += self._peer_tile(peer, src_name)[:rows]

    def sync(self):
        pass
