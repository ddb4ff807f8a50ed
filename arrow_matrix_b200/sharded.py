"""Multi-GPU arrow SpMM: one process per GPU, block-rows of every level sharded across GPUs.

The reference distributes an arrow matrix by giving rank ``i`` the blocks ``A_0i``, ``A_ii``, ``A_i0``
(``arrow/arrow_slim_mpi.py:246-256``), broadcasting the head features ``X_0`` (``:273``), reducing the
head result ``C_0 = sum_i A_0i X_i`` to rank 0 (``:116``) and moving rows between consecutive levels with
all-to-all-v (``arrow/arrow_dec_mpi.py:404-440, 507-550``).  The same algebra here, with a GPU owning a
contiguous *range* of block-rows of every level:

* local matrix of GPU g at a level  = [ rows 0..w of the level restricted to g's columns ]   (partial C_0)
                                       [ g's own block-rows (columns: head + own diagonal)   ]
  in local column numbering ``[head tile | own rows]`` -- ONE SpMM launch per level per GPU;
* ``X_0`` broadcast  -> every GPU copies the head tile out of GPU 0's memory (NVLink peer read);
* ``C_0`` reduce     -> GPU 0 pulls the partial head tiles of its peers and adds them;
* level exchange     -> the owner of a destination row pulls the source row from the peer that owns it
                        (``arrow_gather_rows_multi``: forward with ``to_prev``, backward -- as a gather-add --
                        with ``to_next``); the maps are injective, so there are no atomics and no packing;
* ordering           -> device-side barriers over peer-mapped flags (``arrow_peer_barrier``), no host sync.

``ShardPlan`` is pure numpy (testable anywhere); the engine drives a backend: ``CudaPeerBackend`` (NVLink
peer memory through CUDA IPC) here, a gloo/numpy stand-in in ``tests/`` for the CPU tests.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import decomp


# ------------------------------------------------------------------------------------------------------
# host-side plan
# ------------------------------------------------------------------------------------------------------
class LevelShard:
    """Rank-local view of one level."""
    __slots__ = ("level", "n_blocks", "rows_global", "bounds", "r0", "r1", "own_rows", "hoff", "local_rows",
                 "indptr", "indices", "data", "nnz", "dropped", "fwd_map", "bwd_map",
                 "halo_prev_off", "halo_next_off", "halo_prev_src", "halo_next_src")


class ShardPlan:
    def __init__(self, decomposition: Sequence[Tuple[decomp.Level, np.ndarray]], width: int, rank: int, world: int,
                 block_diagonal: bool = True, n_blocks: Optional[Sequence[int]] = None, partition: str = "locality"):
        """``partition``: 'even' = every level cut into equal contiguous shards; 'locality' (default) = a level whose
        permutation keeps most rows near their partners is cut where its rows map (``decomp.locality_partition``)."""
        self.block_diagonal = bool(block_diagonal)
        self.partition_used = []
        self.width, self.rank, self.world = int(width), int(rank), int(world)
        self.L = len(decomposition)
        self.n_blocks = [decomp.number_of_blocks(B, width) for B, _ in decomposition] if n_blocks is None \
            else [int(b) for b in n_blocks]
        self.perms, self.to_prev, self.to_next, self.sentinel = decomp.prepare_permutations(
            [p for _, p in decomposition], self.n_blocks, width)
        self.levels: List[LevelShard] = []
        w = self.width
        for j, (B, _) in enumerate(decomposition):
            nb = self.n_blocks[j]
            sh = LevelShard()
            sh.level, sh.n_blocks, sh.rows_global = j, nb, nb * w
            bounds_blocks = None
            if partition == "locality" and j > 0 and self.block_diagonal:
                bounds_blocks = decomp.locality_partition(self.to_prev[j], nb, w, self.levels[j - 1].bounds, world)
            self.partition_used.append("even" if bounds_blocks is None else "locality")
            if bounds_blocks is None:
                bounds_blocks = decomp.block_partition(nb, world)
            sh.bounds = bounds_blocks * w                                     # global row bounds per rank
            sh.r0, sh.r1 = int(sh.bounds[rank]), int(sh.bounds[rank + 1])
            sh.own_rows = sh.r1 - sh.r0
            self._layout(sh, rank)
            self._build_local_matrix(sh, B, nb)
            # exchange maps hold GLOBAL rows of the neighbouring level (or -1)
            sh.fwd_map = sh.bwd_map = None
            if j > 0:
                tp = self.to_prev[j][sh.r0:sh.r1].copy()                      # level j row -> level j-1 row
                tp[tp >= self.n_blocks[j - 1] * w] = -1
                sh.fwd_map = tp
            if j < self.L - 1:
                tn = self.to_next[j][sh.r0:sh.r1].copy()                      # level j row -> level j+1 row
                tn[tn >= self.n_blocks[j + 1] * w] = -1
                sh.bwd_map = tn
            self.levels.append(sh)

    # local tile layout of a rank at a level: [head tile (ranks > 0)] [halo: previous block] [own rows] [halo: next block]
    # The halos exist only in the banded layout (blocks A_i,i-1 / A_i,i+1, arrow_mpi.py:211-219) at shard boundaries.
    def _halo_flags(self, level_bounds: np.ndarray, nb: int, g: int):
        w = self.width
        b0, b1 = int(level_bounds[g]) // w, int(level_bounds[g + 1]) // w
        if self.block_diagonal or b1 <= b0:
            return False, False
        return (b0 >= 2), (b1 < nb)

    def hoff_of(self, level: int, g: int) -> int:
        """offset of rank ``g``'s own rows inside its local tile of ``level``"""
        sh = self.levels[level] if level < len(self.levels) else None
        bounds = sh.bounds if sh is not None else None
        nb = self.n_blocks[level]
        if bounds is None:
            bounds = decomp.block_partition(nb, self.world) * self.width
        prev, _ = self._halo_flags(bounds, nb, g)
        return (self.width if g > 0 else 0) + (self.width if prev else 0)

    def _layout(self, sh: LevelShard, rank: int):
        w = self.width
        prev, nxt = self._halo_flags(sh.bounds, sh.n_blocks, rank)
        head = w if rank > 0 else 0
        sh.halo_prev_off = head if prev else -1
        sh.hoff = head + (w if prev else 0)
        sh.halo_next_off = sh.hoff + sh.own_rows if nxt else -1
        sh.local_rows = sh.hoff + sh.own_rows + (w if nxt else 0)
        # where the halo rows live: (owner rank, global first row)
        sh.halo_prev_src = sh.halo_next_src = None
        if prev:
            g0 = sh.r0 - w
            sh.halo_prev_src = (int(np.searchsorted(sh.bounds, g0, side="right") - 1), g0)
        if nxt:
            g1 = sh.r1
            sh.halo_next_src = (int(np.searchsorted(sh.bounds, g1, side="right") - 1), g1)

    def _local_cols(self, sh: LevelShard, idx: np.ndarray) -> np.ndarray:
        """global column -> local column of this rank's tile (-1 if the column is not held locally)"""
        w, rank = self.width, self.rank
        out = np.full(idx.shape, -1, dtype=np.int64)
        own = (idx >= sh.r0) & (idx < sh.r1)
        out[own] = idx[own] - sh.r0 + sh.hoff
        if rank > 0:
            head = idx < w
            out[head] = idx[head]
        if sh.halo_prev_off >= 0:
            lo = sh.r0 - w
            m = (idx >= lo) & (idx < sh.r0)
            out[m] = idx[m] - lo + sh.halo_prev_off
        if sh.halo_next_off >= 0:
            m = (idx >= sh.r1) & (idx < sh.r1 + w)
            out[m] = idx[m] - sh.r1 + sh.halo_next_off
        return out

    def _build_local_matrix(self, sh: LevelShard, B, nb: int):
        w, rank = self.width, self.rank
        parts_ptr, parts_idx, parts_dat = [], [], []
        dropped = 0
        has_data = decomp.level_triplet(B)[0] is not None
        if rank > 0:
            # partial head rows: rows [0, w) of the level restricted to this rank's columns
            ip, idx, dat, d0 = decomp.arrow_rows(B, w, nb, self.block_diagonal, 0, min(w, sh.rows_global))
            dropped += d0 if rank == 0 else 0
            keep = (idx >= sh.r0) & (idx < sh.r1)
            rows = np.repeat(np.arange(ip.size - 1, dtype=np.int64), np.diff(ip))
            cnt = np.bincount(rows[keep], minlength=w)[:w] if keep.any() else np.zeros(w, dtype=np.int64)
            cnt = np.concatenate([cnt, np.zeros(w - cnt.size, dtype=np.int64)]) if cnt.size < w else cnt
            parts_ptr.append(cnt)
            parts_idx.append((idx[keep].astype(np.int64) - sh.r0 + sh.hoff))
            if has_data:
                parts_dat.append(dat[keep])
        if sh.halo_prev_off >= 0:
            parts_ptr.append(np.zeros(w, dtype=np.int64))          # halo rows are inputs only: empty matrix rows
        if sh.own_rows > 0:
            ip, idx, dat, d1 = decomp.arrow_rows(B, w, nb, self.block_diagonal, sh.r0, sh.r1)
            dropped += d1
            idx = idx.astype(np.int64)
            if rank == 0:
                # block-row 0 reaches every column: keep this rank's columns only (peers compute the rest)
                rows = np.repeat(np.arange(ip.size - 1, dtype=np.int64), np.diff(ip))
                local = self._local_cols(sh, idx)
                keep = local >= 0
                # rows of block-row 0 may only use this rank's own columns (halo columns belong to their owner's share)
                keep &= ~((rows < w) & (idx >= sh.r1))
                assert np.all(keep | (rows < w)), "entry outside the arrow / band pattern in an own block-row"
                cnt = np.bincount(rows[keep], minlength=sh.own_rows)
                parts_ptr.append(cnt)
                parts_idx.append(local[keep])
                if has_data:
                    parts_dat.append(dat[keep])
            else:
                local = self._local_cols(sh, idx)
                assert np.all(local >= 0), "entry outside the arrow / band pattern in an own block-row"
                parts_ptr.append(np.diff(ip))
                parts_idx.append(local)
                if has_data:
                    parts_dat.append(dat)
        counts = np.concatenate(parts_ptr) if parts_ptr else np.zeros(0, dtype=np.int64)
        sh.indptr = np.zeros(sh.local_rows + 1, dtype=np.int64)
        if counts.size:
            np.cumsum(counts, out=sh.indptr[1:1 + counts.size])
            sh.indptr[1 + counts.size:] = sh.indptr[counts.size]
        sh.indices = (np.concatenate(parts_idx) if parts_idx else np.zeros(0, np.int64)).astype(np.int32)
        sh.data = np.ascontiguousarray(np.concatenate(parts_dat), dtype=np.float32) if (has_data and parts_dat) else \
            (np.ones(sh.indices.size, dtype=np.float32) if not has_data else np.zeros(0, np.float32))
        sh.nnz = int(sh.indptr[-1])
        sh.dropped = dropped

    def own_bounds(self, level: int) -> np.ndarray:
        return self.levels[level].bounds

    # -- layouts of OTHER ranks (every rank derives the whole routing from the global maps: no set-up traffic) ----
    def layout_of(self, level: int, g: int):
        """``(head_rows, halo_prev_off, hoff, halo_next_off, local_rows, r0, r1)`` of rank ``g`` at ``level``"""
        w = self.width
        bounds = self.levels[level].bounds
        nb = self.n_blocks[level]
        r0, r1 = int(bounds[g]), int(bounds[g + 1])
        prev, nxt = self._halo_flags(bounds, nb, g)
        head = w if g > 0 else 0
        hoff = head + (w if prev else 0)
        own = r1 - r0
        return head, (head if prev else -1), hoff, (hoff + own if nxt else -1), hoff + own + (w if nxt else 0), r0, r1

    def local_to_global(self, level: int, g: int) -> np.ndarray:
        """global row of ``level`` held by every row of rank ``g``'s local tile: [head | halo_prev | own | halo_next]"""
        w = self.width
        head, prev_off, hoff, next_off, local_rows, r0, r1 = self.layout_of(level, g)
        out = np.empty(local_rows, dtype=np.int64)
        if head:
            out[:w] = np.arange(w, dtype=np.int64)
        if prev_off >= 0:
            out[prev_off:prev_off + w] = np.arange(r0 - w, r0, dtype=np.int64)
        out[hoff:hoff + (r1 - r0)] = np.arange(r0, r1, dtype=np.int64)
        if next_off >= 0:
            out[next_off:next_off + w] = np.arange(r1, r1 + w, dtype=np.int64)
        return out

    def local_index(self, level: int, g: int, rows: np.ndarray) -> np.ndarray:
        """position of global ``rows`` of ``level`` inside rank ``g``'s local tile, -1 where ``g`` does not hold the row"""
        w = self.width
        head, prev_off, hoff, next_off, _, r0, r1 = self.layout_of(level, g)
        rows = np.asarray(rows, dtype=np.int64)
        out = np.full(rows.shape, -1, dtype=np.int64)
        if head:
            m = (rows >= 0) & (rows < w)
            out[m] = rows[m]
        if prev_off >= 0:
            m = (rows >= r0 - w) & (rows < r0)
            out[m] = rows[m] - (r0 - w) + prev_off
        if next_off >= 0:
            m = (rows >= r1) & (rows < r1 + w)
            out[m] = rows[m] - r1 + next_off
        m = (rows >= r0) & (rows < r1)
        out[m] = rows[m] - r0 + hoff
        return out

    def composed_maps(self) -> List[Optional[np.ndarray]]:
        """``cmap[j][r]`` = row of level 0 that feeds row ``r`` of level ``j`` (``to_prev`` chained down), -1 when the chain
        leaves the active rows of some level (the reference's sentinel, arrow_dec_mpi.py:740-749)"""
        if getattr(self, "_cmap", None) is None:
            cmap: List[Optional[np.ndarray]] = [None]
            prev = None
            for j in range(1, self.L):
                rows_j, rows_p = self.levels[j].rows_global, self.levels[j - 1].rows_global
                tp = self.to_prev[j][:rows_j]
                valid = tp < rows_p
                safe = np.where(valid, tp, 0)
                cur = np.where(valid, safe if prev is None else prev[safe], -1)
                cmap.append(cur)
                prev = cur
            self._cmap = cmap
        return self._cmap

    def a2a_tables(self, dst_level: int, forward: bool):
        """Pack / unpack tables of one level exchange for an all-to-all-v (the NCCL backend).

        forward: destination = level ``dst_level`` rows, source = level ``dst_level-1`` (``to_prev``);
        backward: destination = level ``dst_level`` rows, source = level ``dst_level+1`` (``to_next``).
        Returns ``pack`` (own source rows in send order), ``send_counts``/``recv_counts`` (rows per peer) and
        ``unpack`` (own destination row -> position in the receive buffer, -1 if not routed).  Same content as
        the reference's ``_all_to_all_tables`` (arrow_dec_mpi.py:325-384), built from the global maps."""
        w = self.width
        src_level = dst_level - 1 if forward else dst_level + 1
        dmap = (self.to_prev[dst_level] if forward else self.to_next[dst_level])[: self.levels[dst_level].rows_global]
        db, sb = self.levels[dst_level].bounds, self.levels[src_level].bounds
        src_rows = self.levels[src_level].rows_global
        dst = np.flatnonzero(dmap < src_rows)                         # routed destination rows (global), ascending
        src = dmap[dst]
        d_rank = np.searchsorted(db, dst, side="right") - 1
        s_rank = np.searchsorted(sb, src, side="right") - 1
        me = self.rank
        # what I send: my source rows, ordered by (destination rank, destination row)
        mine = np.flatnonzero(s_rank == me)
        order = mine[np.lexsort((dst[mine], d_rank[mine]))]
        pack = src[order] - sb[me]
        send_counts = np.bincount(d_rank[order], minlength=self.world).astype(np.int64)
        # what I receive: my destination rows, ordered by (source rank, destination row)
        tome = np.flatnonzero(d_rank == me)
        rorder = tome[np.lexsort((dst[tome], s_rank[tome]))]
        recv_counts = np.bincount(s_rank[rorder], minlength=self.world).astype(np.int64)
        unpack = np.full(self.levels[dst_level].own_rows, -1, dtype=np.int64)
        unpack[dst[rorder] - db[me]] = np.arange(rorder.size, dtype=np.int64)
        return dict(pack=pack.astype(np.int64), send_counts=send_counts, recv_counts=recv_counts, unpack=unpack)


class FusedPlan:
    """Routing of the fused multi-GPU step for one rank (pure numpy, derived from the global maps on every rank).

    Forward: the feature operand of level ``j >= 1`` on this GPU is never materialised.  Column ``c`` of the local level
    matrix stands for the level-``j`` row ``g``; its value is row ``cmap_j[g]`` of level 0.  If this GPU holds that
    level-0 row (own rows, head tile, halo) the column index points into the level-0 tile, otherwise into the
    *receive region*: one slot per remote row, grouped by source GPU, which the owners fill with ONE push kernel
    (``arrow_push_rows``) per step.  ``colmap[j]`` is that re-indexing; columns ``>= x_split`` address the region.

    Backward: ``C_{j-1}[to_prev_j[g]] += C_j[g]`` (arrow_dec_mpi.py:437).  Every row of level ``j-1`` that receives a
    contribution has one slot in its owner's *staging tile*; the level-``j`` SpMM writes each result row straight into
    the slot (``out_which/out_row``: a pointer table, local or over NVLink), and level ``j-1`` adds its staging tile in
    its own epilogue (``add_map``; level 0: one final gather-add).  Head rows (block-row 0 is computed as partial sums
    by every GPU) stay local, are reduced by GPU 0 and delivered by the reduction kernel (``head_which/head_row``).
    """

    def __init__(self, plan: "ShardPlan"):
        self.plan = plan
        pl, me, world, w, L = plan, plan.rank, plan.world, plan.width, plan.L
        cmap = pl.composed_maps()
        b0 = pl.levels[0].bounds
        self.x_split = pl.levels[0].local_rows
        # ---- forward: what every destination needs, in its slot order -----------------------------------------
        self.colmap: List[Optional[np.ndarray]] = [None] * L
        self.ok = True
        per_dest = {}                                                      # d -> (my level-0 tile rows in d's slot order, offset in d's region)
        self.recv_rows = 0
        for d in range(world):
            glob_all, lvl_all, lc_all, src_all = [], [], [], []
            for j in range(1, L):
                glob = pl.local_to_global(j, d)
                inside = glob < pl.levels[j].rows_global
                src0 = np.where(inside, cmap[j][np.where(inside, glob, 0)], -1)
                glob_all.append(glob); src_all.append(src0)
                lvl_all.append(np.full(glob.size, j, dtype=np.int64)); lc_all.append(np.arange(glob.size, dtype=np.int64))
            if not src_all:                                                # a single level: nothing is exchanged
                continue
            src0 = np.concatenate(src_all)
            loc = pl.local_index(0, d, src0)
            remote = (src0 >= 0) & (loc < 0)
            owner = np.searchsorted(b0, np.where(remote, src0, 0), side="right") - 1
            ridx = np.flatnonzero(remote)
            order = ridx[np.argsort(owner[ridx], kind="stable")]          # slot order: (source rank, level, local column)
            counts = np.bincount(owner[order], minlength=world).astype(np.int64)
            offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            if d == me:
                slot = np.full(src0.size, -1, dtype=np.int64)
                slot[order] = np.arange(order.size, dtype=np.int64)
                self.recv_rows = int(order.size)
                self.recv_counts = counts
                pos = 0
                for j in range(1, L):
                    n = src_all[j - 1].size
                    cm = np.where(loc[pos:pos + n] >= 0, loc[pos:pos + n],
                                  np.where(slot[pos:pos + n] >= 0, self.x_split + slot[pos:pos + n], -1))
                    self.colmap[j] = cm
                    idx = pl.levels[j].indices
                    if idx.size and np.any(cm[idx] < 0):
                        self.ok = False                    # a non-zero reads a row behind the sentinel: exchange mode only
                    pos += n
            else:
                mine = order[owner[order] == me]                           # d's slots that I fill, in d's slot order
                if mine.size:
                    per_dest[d] = (pl.local_index(0, me, src0[mine]), int(offs[me]))   # rows + where they start in d's region
        # destination blocks in ROTATED order (me+1, me+2, ...): at any moment GPU r sends to r+t, a perfect matching.  With
        # every list starting at GPU 0 all senders converge on one receiver after each barrier (8 B200: the step took 3.2 ms
        # instead of 1.6 ms; the switch gives each GPU 900 GB/s in AND out, but only if the flows are spread)
        self.push_dest = [d for d in ((me + t) % world for t in range(1, world)) if d in per_dest]
        self.push_src = np.concatenate([per_dest[d][0] for d in self.push_dest]) if self.push_dest else np.zeros(0, dtype=np.int64)
        self.push_bounds = np.concatenate([[0], np.cumsum([per_dest[d][0].size for d in self.push_dest])]).astype(np.int64)
        self.push_off = [per_dest[d][1] for d in self.push_dest]
        assert self.push_src.size == 0 or self.push_src.min() >= 0
        # ---- backward: staging slots of every level that receives --------------------------------------------
        # A staging tile is laid out by SOURCE: [reduced head rows (delivered by GPU 0's reduction) | rows computed by GPU 0 |
        # GPU 1 | ...], each group in ascending row order of the sending level.  A GPU therefore writes the rows it owes a
        # peer into one contiguous block of a local send tile (straight from the SpMM epilogue: local stores) and ships the
        # block with one copy-engine transfer; rows that stay on the GPU go straight into its own staging tile.  (Measured
        # on 2 B200: storing the rows into the peer's memory from the epilogue instead ran the level-1 SpMM at 3.7 ms against
        # 0.75 ms with local stores -- NVLink stores trickling out of a latency-bound kernel reach 170 GB/s.)
        self.stage_rows = [0] * max(L - 1, 0)                              # my staging tile of level j (filled by level j+1)
        self.send_rows = [0] * L                                           # my send tile of level j >= 1
        self.send_plan: List[list] = [[] for _ in range(L)]                # level j: (dest, send_off, rows, dest_stage_off)
        self.recv_plan: List[list] = [[] for _ in range(L)]                # level j: (source, its send_off, rows, my stage_off)
        self.add_map: List[Optional[np.ndarray]] = [None] * L              # local row of level j -> slot in my staging tile
        self.out_which: List[Optional[np.ndarray]] = [None] * L            # level j >= 1: 0 = local C tile, 1 = my staging tile, 2 = my send tile
        self.out_row: List[Optional[np.ndarray]] = [None] * L
        self.head_which: List[Optional[np.ndarray]] = [None] * L           # GPU 0 only: reduced head rows: 1 + d = staging tile of GPU d
        self.head_row: List[Optional[np.ndarray]] = [None] * L
        for j in range(1, L):
            rows_j, rows_p = pl.levels[j].rows_global, pl.levels[j - 1].rows_global
            bp, bj = pl.levels[j - 1].bounds, pl.levels[j].bounds
            hr = min(w, rows_j)
            tp = pl.to_prev[j][:rows_j]
            gs = np.flatnonzero(tp < rows_p)                               # routed rows of level j, ascending
            dest = np.searchsorted(bp, tp[gs], side="right") - 1           # GPU that owns the receiving row
            group = np.where(gs < hr, 0, 1 + (np.searchsorted(bj, gs, side="right") - 1))
            order = np.lexsort((gs, group, dest))                          # slot order inside every destination
            d_sorted, g_sorted, grp_sorted = dest[order], gs[order], group[order]
            d_start = np.searchsorted(d_sorted, np.arange(world + 1), side="left")
            slot = np.arange(order.size, dtype=np.int64) - d_start[d_sorted]
            slot_of = np.full(rows_j, -1, dtype=np.int64)                  # level-j row -> slot at its destination
            slot_of[g_sorted] = slot
            dest_of = np.full(rows_j, -1, dtype=np.int64)
            dest_of[g_sorted] = d_sorted
            self.stage_rows[j - 1] = int(d_start[me + 1] - d_start[me])
            # block (s -> d): rows computed by GPU s that GPU d receives.  counts[s, d]; inside s's send tile the blocks
            # follow each other by destination (its own excluded), inside d's staging tile by source after the head rows
            counts = np.zeros((world, world), dtype=np.int64)
            body = group > 0
            np.add.at(counts, (group[body] - 1, dest[body]), 1)
            heads = np.bincount(dest[~body], minlength=world)
            for s_rank in ((me + t) % world for t in range(1, world)):    # rotated: GPU r pulls from r+t at step t (no hot source)
                if counts[s_rank, me] == 0:
                    continue
                send_off = int(sum(counts[s_rank, d2] for d2 in range(me) if d2 != s_rank))
                stage_off = int(heads[me] + counts[:s_rank, me].sum())
                self.recv_plan[j].append((s_rank, send_off, int(counts[s_rank, me]), stage_off))
            # receiving side: own row of level j-1 -> slot
            shp = pl.levels[j - 1]
            am = np.full(shp.local_rows, -1, dtype=np.int64)
            own = np.arange(shp.r0, shp.r1, dtype=np.int64)
            tn = pl.to_next[j - 1][own]
            has = tn < rows_j
            am[shp.hoff:shp.hoff + own.size] = np.where(has, slot_of[np.where(has, tn, 0)], -1)
            self.add_map[j - 1] = am
            # sending side: my own rows of level j
            sh = pl.levels[j]
            which = np.full(sh.local_rows, -1, dtype=np.int32)
            row = np.zeros(sh.local_rows, dtype=np.int64)
            if me > 0:                                                     # partial head rows stay local (reduced by GPU 0)
                which[:hr] = 0
                row[:hr] = np.arange(hr)
            g = np.arange(sh.r0, sh.r1, dtype=np.int64)
            if g.size:
                d_g, s_g = dest_of[g], slot_of[g]
                part = g < hr                                              # GPU 0's own share of block-row 0: a partial sum too
                remote = (~part) & (d_g >= 0) & (d_g != me)
                # send tile: blocks per destination (ascending), rows in ascending g inside a block = the destination's order
                r_idx = np.flatnonzero(remote)
                r_order = r_idx[np.argsort(d_g[r_idx], kind="stable")]
                send_pos = np.full(g.size, -1, dtype=np.int64)
                send_pos[r_order] = np.arange(r_order.size, dtype=np.int64)
                self.send_rows[j] = int(r_order.size)
                cnt = np.bincount(d_g[r_order], minlength=world) if r_order.size else np.zeros(world, dtype=np.int64)
                off = np.concatenate([[0], np.cumsum(cnt)])
                for d in ((me + t) % world for t in range(1, world)):      # rotated like the forward push
                    if cnt[d] > 0:
                        first = r_order[off[d]]                            # my first row for d: its slot starts my block there
                        self.send_plan[j].append((d, int(off[d]), int(cnt[d]), int(s_g[first])))
                        assert np.array_equal(s_g[r_order[off[d]:off[d + 1]]], s_g[first] + np.arange(cnt[d]))
                wq = np.where(part, 0, np.where(d_g < 0, -1, np.where(d_g == me, 1, 2))).astype(np.int32)
                rq = np.where(part, sh.hoff + (g - sh.r0), np.where(d_g == me, s_g, send_pos))
                which[sh.hoff:sh.hoff + g.size] = wq
                row[sh.hoff:sh.hoff + g.size] = np.where(wq >= 0, rq, 0)
            self.out_which[j], self.out_row[j] = which, row
            if me == 0:
                hg = np.arange(hr, dtype=np.int64)
                self.head_which[j] = np.where(dest_of[hg] >= 0, 1 + dest_of[hg], -1).astype(np.int32)
                self.head_row[j] = np.where(dest_of[hg] >= 0, slot_of[hg], 0)


# ------------------------------------------------------------------------------------------------------
# engine
# ------------------------------------------------------------------------------------------------------
class ShardedArrowEngine:
    """Executes a ShardPlan.  ``backend`` supplies device memory, kernels, peer access and barriers."""

    def __init__(self, plan: ShardPlan, k: int, backend, overlap: bool = False, split_frac: float = 0.5, mode: str = "auto"):
        """``mode``: 'fused' (push exchange folded into the SpMMs, see ``FusedPlan``), 'exchange' (the literal protocol with
        every level's tiles materialised: needed when a non-zero reads a row behind the sentinel, and for
        ``result(level > 0)``), 'auto' = fused when it is exact and the backend can push."""
        if mode not in ("auto", "fused", "exchange"):
            raise ValueError(f"mode must be auto|fused|exchange, got {mode!r}")
        self.plan, self.k, self.be = plan, int(k), backend
        self.overlap = bool(overlap) and hasattr(backend, "side_begin")
        self.overlap_ctas = 3
        self.rank, self.world, self.width, self.L = plan.rank, plan.world, plan.width, plan.L
        be = backend
        # fused step: every rank checks its own level matrices against its routing, all ranks agree
        self.fp = None
        if mode != "exchange" and self.L >= 2 and getattr(backend, "supports_fused", False):
            fp = FusedPlan(plan)
            if int(be.allreduce_sum(0 if fp.ok else 1)) == 0:
                self.fp = fp
        if mode == "fused" and self.fp is None:
            raise ValueError("fused mode requested but a level reads rows behind the sentinel (or the backend cannot push); "
                             "use mode='exchange'")
        self.fused_ok = self.fp is not None
        # overlap=2 (two levels, staged exchange available): level 0 is multiplied in two row parts so that BOTH
        # exchanges hide behind it -- forward behind part a, backward (staged into a buffer) behind part b
        self.split = (self.fp is None and self.overlap and int(overlap) == 2 and plan.L == 2 and plan.world > 1
                      and getattr(backend, "supports_staged_exchange", False))
        self.mode = ("fused-" if self.fp is not None else "exchange-") + type(backend).__name__
        self.graphs = {}
        self.mats, self.fwd, self.bwd = [], [], []
        for sh in plan.levels:
            self.mats.append(be.csr_upload(sh.local_rows, sh.local_rows, sh.indptr, sh.indices, sh.data)
                             if sh.local_rows > 0 else None)
            prev_rows = plan.levels[sh.level - 1].rows_global if sh.level > 0 else 0
            next_rows = plan.levels[sh.level + 1].rows_global if sh.level < self.L - 1 else 0
            self.fwd.append(be.map_upload(sh.fwd_map, prev_rows) if sh.fwd_map is not None else None)
            self.bwd.append(be.map_upload(sh.bwd_map, next_rows) if sh.bwd_map is not None else None)
        # two ping-pong tiles per level, shared with the peers; level 0 gets a second pair for the streaming
        # iteration (upload of step i+1 / download of step i-1 overlap the compute of step i)
        rows_per_level = [max(sh.local_rows, 1) for sh in plan.levels]
        tiles_per_level = [4] + [2] * (self.L - 1)
        if self.split:
            # one more shared "level": the staging tile of the backward exchange into level 0
            self._stage = (self.L, 0)
            rows_per_level.append(max(int(plan.a2a_tables(0, False)["recv_counts"].sum()), 1))
            tiles_per_level.append(1)
            sh0 = plan.levels[0]
            ip = np.asarray(sh0.indptr, dtype=np.int64)
            cut = int(np.searchsorted(ip, ip[0] + split_frac * (ip[-1] - ip[0]), side="left"))
            # part a holds the first `width` local rows: the head block on rank 0, the partial head tile elsewhere --
            # they are reduced (read by / added on rank 0) before part b runs
            self._cut = min(max(cut, min(self.width, sh0.local_rows)), sh0.local_rows)
            self._parts = []
            for r0, r1 in ((0, self._cut), (self._cut, sh0.local_rows)):
                a, b = int(ip[r0]), int(ip[r1])
                self._parts.append(be.csr_upload(r1 - r0, sh0.local_rows, ip[r0:r1 + 1] - a, sh0.indices[a:b], sh0.data[a:b])
                                   if r1 > r0 else None)
        if self.fp is not None:
            # two more kinds of shared tiles: the receive region of the forward push and one staging tile per level
            # that receives contributions from the level below it
            self._recv = (len(rows_per_level), 0)
            rows_per_level.append(max(self.fp.recv_rows, 1))
            tiles_per_level.append(1)
            self._stg, self._snd = [], [None]
            for j in range(self.L - 1):
                self._stg.append((len(rows_per_level), 0))
                rows_per_level.append(max(self.fp.stage_rows[j], 1))
                tiles_per_level.append(1)
            for j in range(1, self.L):
                self._snd.append((len(rows_per_level), 0))
                rows_per_level.append(max(self.fp.send_rows[j], 1))
                tiles_per_level.append(1)
        self.tiles = be.alloc_shared_tiles(rows_per_level, self.k, tiles_per_level=tiles_per_level)
        if self.fp is not None:
            self._setup_fused()
        self.xi = [0] * self.L
        self.ci = [0] * self.L
        self._pair = 0                      # which level-0 pair is active: tile index = 2*pair + {0,1}
        if self.fp is None and hasattr(be, "prepare_exchange"):
            # exchange tables, identity maps, views: everything the step would otherwise create lazily.  Allocation and
            # (above all) cudaFree synchronise the whole device; when the ranks are threads of one process that must not
            # happen while a peer's barrier kernel is already spinning
            be.prepare_exchange(self.L, [min(self.width, sh.rows_global) for sh in plan.levels])
        self.total_nnz_local = sum(sh.nnz for sh in plan.levels)
        self.total_nnz = int(be.allreduce_sum(self.total_nnz_local))
        self.local_rows = plan.levels[0].own_rows
        be.barrier()

    # -- features / results: this rank's own rows of level 0 ------------------------------------------------
    def _other(self, level: int, idx: int) -> int:
        """the partner tile of ``idx`` inside its ping-pong pair"""
        return idx ^ 1

    def set_features(self, X: np.ndarray, sync: bool = True):
        sh = self.plan.levels[0]
        if X.shape != (sh.own_rows, self.k):
            raise ValueError(f"rank {self.rank}: expected features of shape {(sh.own_rows, self.k)}, got {X.shape}")
        if self.xi[0] == self.ci[0]:
            self.xi[0] = self._other(0, self.ci[0])
        self.be.h2d(self.tiles[0][self.xi[0]], sh.hoff, X)
        if sync:
            self.be.sync()

    def rewind_features(self):
        if self.xi[0] == self.ci[0]:
            self.xi[0] = self._other(0, self.ci[0])

    def stream_step(self, X_host: np.ndarray, out_host: np.ndarray):
        """Pipelined host-staged iteration on this rank's rows (see ``ArrowEngine.stream_step``): every rank calls
        it with its own pinned arrays; uploads / step / downloads of consecutive calls overlap."""
        sh = self.plan.levels[0]
        if X_host.shape != (sh.own_rows, self.k) or out_host.shape != (sh.own_rows, self.k):
            raise ValueError(f"rank {self.rank}: expected host arrays of shape {(sh.own_rows, self.k)}")
        ctx = self.be.ctx
        p = self._pair
        self._pair ^= 1
        EV_H2D, EV_MAIN, EV_D2H = 3 * p, 3 * p + 1, 3 * p + 2
        LANE_H2D, LANE_D2H = 1, 2
        x_tile, c_tile = self.tiles[0][2 * p], self.tiles[0][2 * p + 1]
        ctx.event_wait(EV_MAIN, LANE_H2D)
        ctx.h2d_lane(LANE_H2D, x_tile, X_host, row0=sh.hoff)
        ctx.event_record(EV_H2D, LANE_H2D)
        ctx.event_wait(EV_H2D, 0)
        ctx.event_wait(EV_D2H, 0)
        self.xi[0], self.ci[0] = 2 * p, 2 * p + 1
        self.step()
        ctx.event_record(EV_MAIN, 0)
        ctx.event_wait(EV_MAIN, LANE_D2H)
        ctx.d2h_lane(LANE_D2H, self.tiles[0][self.ci[0]], out_host, row0=sh.hoff)
        ctx.event_record(EV_D2H, LANE_D2H)

    def stream_drain(self):
        ctx = self.be.ctx
        ctx.lane_sync(1)
        ctx.sync()
        ctx.lane_sync(2)

    def result(self, level: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        if level > 0 and self.fp is not None:
            raise RuntimeError("levels > 0 are not materialised by the fused step; construct with mode='exchange'")
        sh = self.plan.levels[level]
        return self.be.d2h(self.tiles[level][self.ci[level]], sh.hoff, sh.own_rows, out)

    # -- small uniform API shared with the single-GPU engine ---------------------------------------------------
    @property
    def n_blocks(self):
        return self.plan.n_blocks

    def local_rows_of(self, level: int) -> int:
        return self.plan.levels[level].own_rows

    def zero_rhs(self):
        self.be.barrier()               # no peer may still be reading (direct pulls) or writing (pushes) these tiles
        for pair in self.tiles:
            for t in pair:
                self.be.fill(t, 0.0)
        self.xi = [0] * self.L
        self.ci = [0] * self.L

    def features(self, level: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        sh = self.plan.levels[level]
        return self.be.d2h(self.tiles[level][self.xi[level]], sh.hoff, sh.own_rows, out)

    def spmm_level(self, level: int):
        """One level's distributed arrow product (X_0 broadcast, local SpMM, C_0 reduce): ``B.spmm()``."""
        be, sh = self.be, self.plan.levels[level]
        hr = min(self.width, sh.rows_global)
        be.barrier()
        be.bcast_head((level, self.xi[level]), hr)
        self._halo(level)
        out = self._other(level, self.xi[level])
        if self.mats[level] is not None and sh.local_rows > 0:
            be.spmm(self.mats[level], self.tiles[level][self.xi[level]], self.tiles[level][out])
        self.ci[level] = out
        be.barrier()
        be.reduce_head((level, out), hr)
        be.barrier()

    def ensure_level_tiles(self):
        """Per-level tiles exist in both modes; the fused step just never fills levels > 0 -- callers that want them
        (``result(level > 0)``, the phase-by-phase ``_propagate_features`` / ``_aggregate`` surface) get the literal protocol."""
        if self.fp is not None:
            self.fp = None
            self.mode = "exchange-" + type(self.be).__name__

    def sync(self):
        if getattr(self.be, "ctx", None) is not None:
            self.stream_drain()             # copy lanes + main stream
        else:
            self.be.sync()

    def close(self):
        if getattr(self.be, "ctx", None) is not None:
            if hasattr(self.be, "close"):
                self.be.close()
            else:
                self.be.sync()
                self.be.ctx.close()

    # -- the iteration -----------------------------------------------------------------------------------------
    def propagate_features(self):
        be, pl = self.be, self.plan
        be.barrier()                                            # every rank's level-0 features are in place
        for j in range(1, self.L):
            sh, prev = pl.levels[j], pl.levels[j - 1]
            # C_i[perm] = recvbuf (arrow_dec_mpi.py:544): each routed row comes from the GPU that owns it
            be.pull_rows(dst=(j, self.ci[j]), dst_off=sh.hoff, src=(j - 1, self.xi[j - 1]), src_bounds=prev.bounds,
                         row_map=self.fwd[j], accumulate=False, forward=True)
            self.xi[j] = self.ci[j]                             # set_features(C_i) (:545)
            be.barrier()
        # X_0 broadcast of every level (arrow_slim_mpi.py:273) and, in the banded layout, the halo tiles
        for j in range(self.L):
            be.bcast_head((j, self.xi[j]), min(self.width, pl.levels[j].rows_global))
            self._halo(j)

    def spmm(self):
        for j in range(self.L):
            out = self._other(j, self.xi[j])
            if self.mats[j] is not None and self.plan.levels[j].local_rows > 0:
                self.be.spmm(self.mats[j], self.tiles[j][self.xi[j]], self.tiles[j][out])
            self.ci[j] = out

    def aggregate(self):
        be, pl = self.be, self.plan
        be.barrier()                                            # all partial head tiles are written
        # C_0 = sum_i A_0i X_i (Reduce to rank 0, arrow_slim_mpi.py:116)
        for j in range(self.L):
            be.reduce_head((j, self.ci[j]), min(self.width, pl.levels[j].rows_global))
        be.barrier()
        for j in range(self.L - 1, 0, -1):
            prev, sh = pl.levels[j - 1], pl.levels[j]
            # C_{j-1}[to_prev[r]] += C_j[r] as a gather-add over this rank's level j-1 rows (:437)
            be.pull_rows(dst=(j - 1, self.ci[j - 1]), dst_off=prev.hoff, src=(j, self.ci[j]), src_bounds=sh.bounds,
                         row_map=self.bwd[j - 1], accumulate=True, forward=False)
            self.xi[j - 1] = self.ci[j - 1]                     # set_features(C_i) (:438)
            if j > 1:
                be.barrier()

    def _halo(self, j: int):
        """banded layout: fetch the neighbouring blocks' feature rows that sit on other GPUs (the reference's
        neighbour tile exchange, arrow_mpi.py:150-162)"""
        sh = self.plan.levels[j]
        w = self.width
        for off, src in ((sh.halo_prev_off, sh.halo_prev_src), (sh.halo_next_off, sh.halo_next_src)):
            if off is None or off < 0 or src is None:
                continue
            g, first_row = src
            peer_sh_r0 = int(sh.bounds[g])
            self.be.copy_rows_from_peer(dst=(j, self.xi[j]), dst_off=off, peer=g, src=(j, self.xi[j]),
                                        src_off=self.plan.hoff_of(j, g) + first_row - peer_sh_r0, rows=w)

    def _spmm_one(self, j: int):
        out = self._other(j, self.xi[j])
        if self.mats[j] is not None and self.plan.levels[j].local_rows > 0:
            self.be.spmm(self.mats[j], self.tiles[j][self.xi[j]], self.tiles[j][out])
        self.ci[j] = out

    def _spmm_part(self, part: int, out: int):
        """rows [0, cut) (part 0) or [cut, local_rows) (part 1) of this rank's level-0 product"""
        A = self._parts[part]
        if A is None:
            return
        r0, r1 = ((0, self._cut), (self._cut, self.plan.levels[0].local_rows))[part]
        self.be.spmm(A, self.tiles[0][self.xi[0]], self.be.tile_view(0, out, r0, r1 - r0))

    def _step_split(self):
        """two levels: forward exchange || level-0 part a ; level 1 ; head reduce ; staged backward exchange ||
        level-0 part b ; local add of the staged rows"""
        be, pl = self.be, self.plan
        sh0, sh1 = pl.levels
        hr0, hr1 = min(self.width, sh0.rows_global), min(self.width, sh1.rows_global)
        be.barrier()                                            # every rank's level-0 features are in place
        be.side_begin()
        be.pull_rows(dst=(1, self.ci[1]), dst_off=sh1.hoff, src=(0, self.xi[0]), src_bounds=sh0.bounds,
                     row_map=self.fwd[1], accumulate=False, forward=True, side=True)
        self.xi[1] = self.ci[1]
        be.barrier(side=True)
        be.bcast_head((0, self.xi[0]), hr0)
        self._halo(0)
        out0 = self._other(0, self.xi[0])
        be.limit_spmm(self.overlap_ctas)
        self._spmm_part(0, out0)
        be.limit_spmm(0)
        be.side_join()
        be.bcast_head((1, self.xi[1]), hr1)
        self._halo(1)
        self._spmm_one(1)
        self.ci[0] = out0
        be.barrier()                                            # all partial head tiles are written (part a holds them)
        be.reduce_head((0, self.ci[0]), hr0)
        be.reduce_head((1, self.ci[1]), hr1)
        be.barrier()
        be.side_begin()
        be.stage_rows(dst_level=0, src=(1, self.ci[1]), forward=False, stage=self._stage, side=True)
        be.limit_spmm(self.overlap_ctas)
        self._spmm_part(1, out0)
        be.limit_spmm(0)
        be.side_join()
        # C_0[to_prev[r]] += C_1[r] (arrow_dec_mpi.py:437): the staged rows are local now
        be.apply_staged(dst=(0, self.ci[0]), dst_off=sh0.hoff, dst_level=0, forward=False, stage=self._stage, accumulate=True)
        self.xi[0] = self.ci[0]                                 # set_features(C_i) (:438)

    # -- fused step ----------------------------------------------------------------------------------------------
    def _setup_fused(self):
        be, pl, fp = self.be, self.plan, self.fp
        n_cols = fp.x_split + max(fp.recv_rows, 1)
        self.f_mats, self.f_tables, self.f_head_tables, self.f_add = [None] * self.L, [None] * self.L, [None] * self.L, [None] * self.L
        me = self.rank
        for j in range(1, self.L):
            if self.mats[j] is not None:
                self.f_mats[j] = be.fused_matrix(self.mats[j], fp.colmap[j], n_cols)
            # every SpMM store is local: the level's own tile (partial head rows), my staging tile, my send tile
            self.f_tables[j] = be.out_table([(me,) + (j, 0), (me,) + self._stg[j - 1], (me,) + self._snd[j]],
                                            fp.out_which[j], fp.out_row[j])
            if me == 0:
                # the reduced head rows are few (one block-row): GPU 0 stores them straight into the owners' staging tiles
                self.f_head_tables[j] = be.out_table([(me,) + (j, 0)] + [(g,) + self._stg[j - 1] for g in range(self.world)],
                                                     fp.head_which[j], fp.head_row[j])
        for j in range(self.L - 1):
            n = pl.levels[j].own_rows if j == 0 else pl.levels[j].local_rows
            am = fp.add_map[j][pl.levels[j].hoff:pl.levels[j].hoff + n] if j == 0 else fp.add_map[j]
            self.f_add[j] = be.map_upload(am, max(fp.stage_rows[j], 1))
        self.f_push = be.push_plan(self._recv, fp.push_src, fp.push_bounds, fp.push_off, fp.push_dest, pl.levels[0].local_rows)
        # backward exchange, variant "push": the same kernel ships the blocks of my send tile into the peers' staging tiles
        self.f_bpush = [None] * self.L
        for j in range(1, self.L):
            plan_j = fp.send_plan[j]
            src = np.concatenate([np.arange(o, o + r, dtype=np.int64) for _, o, r, _ in plan_j]) if plan_j else np.zeros(0, dtype=np.int64)
            bounds = np.concatenate([[0], np.cumsum([r for _, _, r, _ in plan_j])]).astype(np.int64)
            self.f_bpush[j] = be.push_plan(self._stg[j - 1], src, bounds, [so for _, _, _, so in plan_j], [d for d, _, _, _ in plan_j],
                                           max(fp.send_rows[j], 1))
        self.bwd_mode = "push"           # "push": SM kernel, all peers at once | "pull": one copy-engine transfer per peer
        self.side_ctas, self.main_ctas = 2, 2

    def _step_fused(self, dry: bool = False):
        """forward push || level-0 SpMM ; deepest level first: SpMM with the [tile | receive region] operand, rows written
        into my staging / send tiles, one copy-engine transfer per peer ; head reductions ; one final gather-add.  ``dry``: the
        same kernels with every cross-GPU effect removed (no push, no copies, no barriers, no head reductions) -- the time this
        step would take if communication were free; bench.py reports the difference as exposed communication."""
        be, pl = self.be, self.plan
        side = self.overlap
        L = self.L
        x = (0, self.xi[0])
        out0 = self._other(0, self.xi[0])
        hr = [min(self.width, pl.levels[j].rows_global) for j in range(L)]
        if not dry:
            be.barrier()                                        # every rank's level-0 features are in place
            be.bcast_head(x, hr[0])
            self._halo(0)
        if side:
            be.side_begin()
        if not dry:
            be.push(self.f_push, x, side=side)                  # forward exchange: one pass, NVLink stores
            be.barrier(side)
        if side:
            be.limit_spmm(self.side_ctas)
        for j in range(L - 1, 0, -1):
            if self.f_mats[j] is not None and pl.levels[j].local_rows > 0:
                be.spmm_fused(self.f_mats[j], x, self._recv, self.fp.x_split, self.f_tables[j],
                              add=self._stg[j] if j < L - 1 else None, add_map=self.f_add[j] if j < L - 1 else None, side=side)
            if not dry:
                if self.bwd_mode == "push":
                    be.push(self.f_bpush[j], self._snd[j], side=side)   # my blocks into the peers' staging tiles, all peers at once
                be.barrier(side)                                # every peer's send tile / pushed block and partial head rows are written
                if self.bwd_mode != "push":
                    # one contiguous block per peer, pulled by the copy engine (no SM time; one source at a time)
                    for src_rank, src_off, rows, dst_off in self.fp.recv_plan[j]:
                        be.copy_rows_from_peer(dst=self._stg[j - 1], dst_off=dst_off, peer=src_rank, src=self._snd[j], src_off=src_off,
                                               rows=rows, side=side)
                if self.rank == 0:
                    be.reduce_rows((j, 0), hr[j], table=self.f_head_tables[j], side=side)
                be.barrier(side)                                # the reduced head rows have landed; send tiles may be rewritten
        if side:
            be.limit_spmm(self.main_ctas)
        if self.mats[0] is not None and pl.levels[0].local_rows > 0:
            be.spmm(self.mats[0], self.tiles[0][x[1]], self.tiles[0][out0])
        be.limit_spmm(0)
        if not dry:
            be.barrier()                                        # all partial head tiles of level 0 are written
            if self.rank == 0:
                be.reduce_rows((0, out0), hr[0], table=None)
        if side:
            be.side_join()
        # C_0[to_prev[r]] += C_1[r] (arrow_dec_mpi.py:437): the staged rows are local now
        be.final_add((0, out0), pl.levels[0].hoff, pl.levels[0].own_rows, self._stg[0], self.f_add[0])
        self.ci[0] = out0
        self.xi[0] = out0                                       # set_features(C_i) (:438)

    def _step_graph(self):
        """The fused step as ONE host call: recorded once per ping-pong parity (CUDA graph, both lanes, barriers
        included -- their epochs live in device memory) after one plain run that settles every lazy allocation."""
        ctx = self.be.ctx
        key = self.xi[0]
        g = self.graphs.get(key)
        if g is None:
            warm = self.__dict__.setdefault("_graph_warm", set())
            if key not in warm:
                warm.add(key)
                return self._step_fused()
            xi, ci = list(self.xi), list(self.ci)
            ctx.graph_begin()
            try:
                self._step_fused()                      # recorded, not executed
            finally:
                g = ctx.graph_end()
            self.graphs[key] = g
            self.xi, self.ci = xi, ci
            # instantiating a graph allocates (and may synchronise the device); with rank threads sharing one device
            # context that must not overlap a peer's replay, whose first kernel is a spinning barrier: every rank
            # finishes its instantiation before anyone launches
            self.be.comm.Barrier()
        ctx.graph_launch(g)
        out0 = self._other(0, self.xi[0])
        self.ci[0] = out0
        self.xi[0] = out0

    def step(self):
        if self.fp is not None:
            if getattr(self, "use_graphs", False) and getattr(self.be, "ctx", None) is not None:
                return self._step_graph()
            return self._step_fused()
        if self.split:
            return self._step_split()
        if not self.overlap:
            self.propagate_features()
            self.spmm()
            self.aggregate()
            return
        # overlap: the forward exchange (NVLink pulls + their barriers) runs on a side lane while the main lane
        # broadcasts level 0's head tile and multiplies level 0 -- both only read level-0 features
        be, pl = self.be, self.plan
        be.barrier()
        be.side_begin()
        for j in range(1, self.L):
            sh, prev = pl.levels[j], pl.levels[j - 1]
            be.pull_rows(dst=(j, self.ci[j]), dst_off=sh.hoff, src=(j - 1, self.xi[j - 1]), src_bounds=prev.bounds,
                         row_map=self.fwd[j], accumulate=False, forward=True, side=True)
            self.xi[j] = self.ci[j]
            be.barrier(side=True)
        be.bcast_head((0, self.xi[0]), min(self.width, pl.levels[0].rows_global))
        self._halo(0)
        be.limit_spmm(self.overlap_ctas)        # leave SM resources to the exchange kernels on the side lane
        self._spmm_one(0)
        be.limit_spmm(0)
        be.side_join()
        for j in range(1, self.L):
            be.bcast_head((j, self.xi[j]), min(self.width, pl.levels[j].rows_global))
            self._halo(j)
            self._spmm_one(j)
        self.aggregate()

    # -- accounting ----------------------------------------------------------------------------------------------
    def time_level_spmm(self, j: int, iters: int, warmup: int = 3) -> float:
        """Average duration (ms) of this rank's level-``j`` SpMM launch alone (CUDA events on its stream)."""
        ctx = self.be.ctx
        sh = self.plan.levels[j]
        if self.mats[j] is None or sh.local_rows == 0:
            return 0.0
        src = self.tiles[j][self.xi[j]]
        scratch = ctx.dense_alloc(sh.local_rows, self.k)
        for _ in range(warmup):
            ctx.spmm(self.mats[j], src, scratch)
        ctx.timer_start(5)
        for _ in range(iters):
            ctx.spmm(self.mats[j], src, scratch)
        ctx.timer_stop(5)
        ms = ctx.timer_ms(5) / iters
        scratch.free()
        return ms

    def level_bytes(self, j: int) -> float:
        """algorithmic bytes of this rank's level-``j`` launch"""
        sh = self.plan.levels[j]
        return sh.nnz * 8 + (sh.local_rows + 1) * 4 + 2.0 * sh.local_rows * self.k * 4

    def flops_per_step(self) -> float:
        return 2.0 * self.total_nnz * self.k

    def algorithmic_bytes_per_step(self) -> float:
        total = 0.0
        pl = self.plan
        for j in range(self.L):
            rows = pl.levels[j].rows_global
            total += (rows + 1) * 4 + 2.0 * rows * self.k * 4
            if j > 0:
                m = int(np.count_nonzero(pl.to_prev[j][:rows] < pl.levels[j - 1].rows_global))
                total += 5.0 * m * self.k * 4
        return total + 8.0 * self.total_nnz

    @property
    def ctx(self):
        return self.be.ctx


# ------------------------------------------------------------------------------------------------------
# CUDA backend: NVLink peer memory through CUDA IPC
# ------------------------------------------------------------------------------------------------------
class CudaPeerBackend:
    def __init__(self, comm, device: int, width: int, stream: Optional[int] = None, plan: Optional[ShardPlan] = None):
        """``plan`` given  -> packed exchange (default): rows are packed locally in the receiver's row order and the
        receiver reads its peers' contiguous regions (sequential NVLink reads).  ``plan=None`` -> direct pulls of
        individual rows at random peer addresses; measured on B200: fine while the peer range stays under ~1 GB
        (669 GB/s), collapsing to ~100 GB/s at 2.5 GB (TLB misses on the NVLink path)."""
        from . import _lib
        self.width = int(width)
        self.plan = plan
        self._xtab = {}
        self._lib = _lib
        self.comm = comm
        self.rank, self.world = comm.Get_rank(), comm.Get_size()
        self.ctx = _lib.Context(device, stream)
        self._preloaded = set()
        if os.environ.get("ARROW_DEBUG_SYNC") == "1":
            self._wrap_debug_sync()
        self._tiles = None
        self._peer = None         # [rank][level][which] -> Dense (imported or own)
        self._flags = None
        self._ident = {}

    def _wrap_debug_sync(self):
        """ARROW_DEBUG_SYNC=1: synchronise after every backend operation so that an asynchronous CUDA error is reported by
        the operation that caused it (debugging aid; serialises the lanes)"""
        import functools
        for name in ("spmm", "barrier", "pull_rows", "bcast_head", "reduce_head", "copy_rows_from_peer", "push", "spmm_fused",
                     "reduce_rows", "final_add", "stage_rows", "apply_staged"):
            fn = getattr(self, name)

            def wrapped(*a, _fn=fn, _name=name, **kw):
                out = _fn(*a, **kw)
                try:
                    for lane in (3, 0):
                        self.ctx.lane_sync(lane)
                except Exception as e:      # noqa: BLE001
                    raise RuntimeError(f"rank {self.rank}: CUDA error surfaced right after backend.{_name}{a}: {e}") from e
                return out
            setattr(self, name, functools.wraps(fn)(wrapped))

    def close(self):
        """Collective.  An exported arena must outlive every peer's mapping of it (CUDA IPC): all ranks finish their
        work and drop their mappings, meet, and only then free their own memory."""
        if self.ctx is None:
            return
        ctx = self.ctx
        try:
            for lane in (1, 2, 3):
                ctx.lane_sync(lane)
            ctx.sync()
        except Exception:       # noqa: BLE001  (a poisoned context still has to release its mappings)
            pass
        for g, arena in enumerate(getattr(self, "_arenas", []) or []):
            if g != self.rank:
                try:
                    arena.free()
                except Exception:       # noqa: BLE001
                    pass
        try:
            self.comm.Barrier()
        finally:
            ctx.close()
            self.ctx = None

    def csr_upload(self, n_rows, n_cols, indptr, indices, data):
        return self.ctx.csr_upload(n_rows, n_cols, indptr, indices, data)

    def map_upload(self, m, limit):
        return self.ctx.map_upload(m, limit)

    def alloc_shared_tiles(self, rows_per_level: Sequence[int], k: int, tiles_per_level: Optional[Sequence[int]] = None):
        """One arena per rank (a single cudaMalloc => a single IPC handle): ping-pong tiles per level + flags."""
        ctx = self.ctx
        self.k = k
        if k not in self._preloaded:
            # no kernel may be loaded for the first time while a peer barrier spins (lazy module loading synchronises
            # the context): load everything this feature width can launch now
            ctx.preload_kernels(k)
            self._preloaded.add(k)
        tiles_per_level = list(tiles_per_level) if tiles_per_level is not None else [2] * len(rows_per_level)
        align = 64                                              # floats (256 bytes)
        offs, pos = [], 128                                     # first 128 floats: barrier flags of the two lanes
        for r, nt in zip(rows_per_level, tiles_per_level):
            pair = []
            for _ in range(nt):
                pair.append(pos)
                pos += -(-(r * k) // align) * align
            offs.append(pair)
        self._send_off, self._send_rows = pos, 0
        if self.plan is not None:
            # one shared send buffer (exchanges are separated by barriers): the largest pack of any exchange
            for lvl in range(self.plan.L):
                for fwd in (True, False):
                    if (fwd and lvl == 0) or (not fwd and lvl == self.plan.L - 1):
                        continue
                    self._send_rows = max(self._send_rows, int(self.plan.a2a_tables(lvl, fwd)["send_counts"].sum()))
            pos += -(-(max(self._send_rows, 1) * k) // align) * align
        arena_rows = max(-(-pos // 64), (2 << 20) // 256)       # >= 2 MiB so the driver gives it its own block
        self._arena = ctx.dense_alloc(arena_rows, 64)
        ctx.sync()
        import os
        mine = dict(handle=self._arena.ipc_export(), arena_rows=arena_rows, offs=offs, rows=list(rows_per_level),
                    send_off=self._send_off, send_rows=self._send_rows, pid=os.getpid(), ptr=self._arena.device_ptr(),
                    device=ctx.device)
        everyone = self.comm.allgather(mine)
        self._peer, self._flags, self._flags_side, self._arenas, self._arena_base = [], [], [], [], []
        for g, info in enumerate(everyone):
            if g == self.rank:
                arena = self._arena
            elif info["pid"] == os.getpid():
                # ranks as threads of one process (comm.ThreadComm): same address space, no IPC mapping
                if info["device"] != ctx.device:
                    raise NotImplementedError("in-process ranks on different devices need peer access; use one process per GPU")
                arena = ctx.dense_wrap(info["ptr"], info["arena_rows"], 64)
            else:
                arena = ctx.ipc_import(info["handle"], info["arena_rows"], 64)
            self._arenas.append(arena)
            base = arena.device_ptr()
            self._arena_base.append((base, info["send_off"]))
            self._flags.append(ctx.dense_wrap(base, 1, 64))
            self._flags_side.append(ctx.dense_wrap(base + 64 * 4, 1, 64))
            self._peer.append([[ctx.dense_wrap(base + o * 4, r, k) for o in pair] for pair, r in zip(info["offs"], info["rows"])])
        self._tiles = self._peer[self.rank]
        self._views = {}
        return self._tiles

    def _view(self, g: int, level: int, which: int, off: int, rows: int):
        """Dense handle on rows [off, off+rows) of a (peer) tile."""
        key = (g, level, which, off, rows)
        v = self._views.get(key)
        if v is None:
            base = self._peer[g][level][which]
            v = self.ctx.dense_wrap(base.device_ptr() + off * self.k * 4, rows, self.k)
            self._views[key] = v
        return v

    def h2d(self, tile, off, X):
        tile.h2d(X, row0=off)

    def fill(self, tile, v):
        tile.fill(v)

    def d2h(self, tile, off, rows, out=None):
        return tile.d2h(out, row0=off, rows=rows)

    def sync(self):
        self.ctx.sync()

    SIDE = 3            # lane id of the side stream (lanes 1 and 2 are the host copy lanes)

    def barrier(self, side: bool = False):
        if self.world <= 1:
            return
        if side:
            self.ctx.set_lane(self.SIDE)
            self.ctx.peer_barrier(self._flags_side, self.rank)
            self.ctx.set_lane(0)
        else:
            self.ctx.peer_barrier(self._flags, self.rank)

    def limit_spmm(self, ctas_per_sm: int):
        self.ctx.set_option(self.ctx.OPT_SPMM_CTAS_PER_SM, ctas_per_sm)

    def side_begin(self):
        """the side lane waits for everything issued so far on the main lane"""
        self.ctx.lane_wait(self.SIDE, 0)

    def side_join(self):
        """the main lane waits for everything issued so far on the side lane"""
        self.ctx.lane_wait(0, self.SIDE)

    def allreduce_sum(self, v):
        return sum(self.comm.allgather(int(v)))

    def spmm(self, A, X, C):
        self.ctx.spmm(A, X, C)

    def pull_rows(self, dst, dst_off, src, src_bounds, row_map, accumulate, forward=True, side=False):
        if self.plan is not None:
            return self._packed_exchange(dst, dst_off, src, accumulate, forward, side)
        lvl, which = dst
        n = row_map.n
        if n == 0:
            return
        if side:
            self.ctx.set_lane(self.SIDE)
        try:
            self._pull_rows(dst, dst_off, src, src_bounds, row_map, accumulate)
        finally:
            if side:
                self.ctx.set_lane(0)

    def _pull_rows(self, dst, dst_off, src, src_bounds, row_map, accumulate):
        lvl, which = dst
        n = row_map.n
        d = self._view(self.rank, lvl, which, dst_off, n)
        srcs = []
        for g in range(self.world):
            own = int(src_bounds[g + 1] - src_bounds[g])
            hoff = self._hoff(g, src[0])
            srcs.append(self._view(g, src[0], src[1], hoff if own > 0 else 0, max(own, 0)))
        self.ctx.gather_rows_multi(d, srcs, [int(b) for b in src_bounds], row_map, accumulate=accumulate)

    # -- packed exchange -------------------------------------------------------------------------------------
    def _exchange_table(self, dst_level: int, forward: bool):
        key = (dst_level, forward)
        t = self._xtab.get(key)
        if t is None:
            raw = self.plan.a2a_tables(dst_level, forward)
            counts = self.comm.allgather([int(c) for c in raw["send_counts"]])       # counts[s][d]
            src_level = dst_level - 1 if forward else dst_level + 1
            n_send, n_recv = int(raw["send_counts"].sum()), int(raw["recv_counts"].sum())
            # where, inside peer s's send buffer, the rows meant for me start
            region = [sum(counts[s][:self.rank]) for s in range(self.world)]
            rb = np.concatenate([[0], np.cumsum(raw["recv_counts"])]).astype(np.int64)
            t = dict(n_send=n_send, n_recv=n_recv, region=region, recv_bounds=[int(b) for b in rb],
                     recv_counts=[int(c) for c in raw["recv_counts"]],
                     pack=self.ctx.map_upload(raw["pack"], max(self.plan.levels[src_level].own_rows, 1)),
                     unpack=self.ctx.map_upload(raw["unpack"], max(n_recv, 1)))
            self._xtab[key] = t
        return t

    def prepare_exchange(self, L: int, head_rows):
        if self.plan is not None:
            for lvl in range(L):
                if lvl > 0:
                    self._exchange_table(lvl, True)
                if lvl < L - 1:
                    self._exchange_table(lvl, False)
        if self.rank == 0:
            for rows in set(int(r) for r in head_rows):
                if rows not in self._ident:
                    self._ident[rows] = self.ctx.map_upload(np.arange(rows, dtype=np.int64), rows)
        self.ctx.sync()

    def _raw_view(self, g: int, float_off: int, rows: int):
        key = ("raw", g, float_off, rows)
        v = self._views.get(key)
        if v is None:
            v = self.ctx.dense_wrap(self._arena_base[g][0] + float_off * 4, rows, self.k)
            self._views[key] = v
        return v

    def _packed_exchange(self, dst, dst_off, src, accumulate, forward, side):
        t = self._exchange_table(dst[0], forward)
        src_sh = self.plan.levels[src[0]]
        if side:
            self.ctx.set_lane(self.SIDE)
        try:
            if t["n_send"] > 0:        # pack my source rows, grouped by destination rank, in the destination's row order
                self.ctx.gather_rows(self._raw_view(self.rank, self._send_off, t["n_send"]),
                                     self._view(self.rank, src[0], src[1], src_sh.hoff, src_sh.own_rows), t["pack"])
        finally:
            if side:
                self.ctx.set_lane(0)
        self.barrier(side)                 # every peer's pack is complete
        n = self.plan.levels[dst[0]].own_rows
        if n > 0 and t["n_recv"] > 0:
            if side:
                self.ctx.set_lane(self.SIDE)
            try:
                srcs = [self._raw_view(g, self._arena_base[g][1] + t["region"][g] * self.k, t["recv_counts"][g])
                        for g in range(self.world)]
                self.ctx.gather_rows_multi(self._view(self.rank, dst[0], dst[1], dst_off, n), srcs, t["recv_bounds"],
                                           t["unpack"], accumulate=accumulate)
            finally:
                if side:
                    self.ctx.set_lane(0)

    @property
    def supports_staged_exchange(self) -> bool:
        return self.plan is not None and type(self) is CudaPeerBackend

    # -- fused step primitives (see FusedPlan) ----------------------------------------------------------------------
    @property
    def supports_fused(self) -> bool:
        return type(self) is CudaPeerBackend

    def _lane(self, side: bool):
        self.ctx.set_lane(self.SIDE if side else 0)

    def fused_matrix(self, A, colmap, n_cols):
        return A.remap_columns(self.ctx.map_upload(colmap, n_cols), n_cols)

    def out_table(self, tiles, which, row):
        """pointer table over ``tiles`` = [(rank, level, index), ...]"""
        return self.ctx.ptrtable_upload([self._peer[g][lv][ix] for g, lv, ix in tiles], which, row)

    def push_plan(self, recv, src_rows, bounds, offs, dests, src_limit):
        """block i = rows bounds[i]..bounds[i+1] of ``src_rows`` -> rows offs[i].. of GPU dests[i]'s receive region"""
        m = self.ctx.map_upload(src_rows, max(int(src_limit), 1))
        dsts = [self._view(d, recv[0], recv[1], int(offs[i]), int(bounds[i + 1] - bounds[i])) for i, d in enumerate(dests)]
        return dict(map=m, dsts=dsts, bounds=[int(b) for b in bounds], n=int(bounds[-1]) if len(dests) else 0)

    def push(self, pp, x, side=False):
        if pp["n"] == 0:
            return
        self._lane(side)
        try:
            self.ctx.push_rows(pp["dsts"], pp["bounds"], self._tiles[x[0]][x[1]], pp["map"])
        finally:
            self._lane(False)

    def spmm_fused(self, A, x, recv, x_split, table, add=None, add_map=None, side=False):
        self._lane(side)
        try:
            self.ctx.spmm_ex(A, self._tiles[x[0]][x[1]], X2=self._tiles[recv[0]][recv[1]], x_split=x_split, out_table=table,
                             add=self._tiles[add[0]][add[1]] if add is not None else None, add_map=add_map)
        finally:
            self._lane(False)

    def reduce_rows(self, tile, rows, table=None, side=False):
        """sum of every rank's first ``rows`` rows of ``tile`` (the partial head tiles), in rank order: into my own tile,
        or -- with a table -- wherever each row is routed"""
        self._lane(side)
        try:
            srcs = [self._view(g, tile[0], tile[1], 0, rows) for g in range(self.world)]
            self.ctx.reduce_rows(srcs, rows, dst=None if table is not None else srcs[self.rank], out_table=table)
        finally:
            self._lane(False)

    def final_add(self, dst, dst_off, rows, stage, add_map):
        if rows > 0:
            self.ctx.gather_rows(self._view(self.rank, dst[0], dst[1], dst_off, rows), self._tiles[stage[0]][stage[1]],
                                 add_map, accumulate=True)

    def tile_view(self, level: int, which: int, off: int, rows: int):
        return self._view(self.rank, level, which, off, rows)

    def stage_rows(self, dst_level: int, src, forward: bool, stage, side: bool = False):
        """First half of the packed exchange: pack -> barrier -> copy this rank's region of every peer's send tile,
        back to back, into the local staging tile (plain sequential NVLink reads; nothing is added yet)."""
        t = self._exchange_table(dst_level, forward)
        src_sh = self.plan.levels[src[0]]
        if side:
            self.ctx.set_lane(self.SIDE)
        try:
            if t["n_send"] > 0:
                self.ctx.gather_rows(self._raw_view(self.rank, self._send_off, t["n_send"]),
                                     self._view(self.rank, src[0], src[1], src_sh.hoff, src_sh.own_rows), t["pack"])
        finally:
            if side:
                self.ctx.set_lane(0)
        self.barrier(side)                 # every peer's pack is complete
        if t["n_recv"] > 0:
            if side:
                self.ctx.set_lane(self.SIDE)
            try:
                pos = 0
                for g in range(self.world):
                    c = t["recv_counts"][g]
                    if c > 0:
                        self._view(self.rank, stage[0], stage[1], pos, c).copy_from(
                            self._raw_view(g, self._arena_base[g][1] + t["region"][g] * self.k, c), rows=c)
                    pos += c
            finally:
                if side:
                    self.ctx.set_lane(0)

    def apply_staged(self, dst, dst_off: int, dst_level: int, forward: bool, stage, accumulate: bool):
        """Second half: scatter / add the staged rows into place (local HBM only)."""
        t = self._exchange_table(dst_level, forward)
        n = self.plan.levels[dst[0]].own_rows
        if n > 0 and t["n_recv"] > 0:
            self.ctx.gather_rows(self._view(self.rank, dst[0], dst[1], dst_off, n),
                                 self._view(self.rank, stage[0], stage[1], 0, t["n_recv"]), t["unpack"], accumulate=accumulate)

    def _hoff(self, g: int, level: int = 0) -> int:
        plan = self.plan if self.plan is not None else getattr(self, "layout_plan", None)
        if plan is not None:
            return plan.hoff_of(level, g)
        return self.width if g > 0 else 0

    def copy_rows_from_peer(self, dst, dst_off, peer, src, src_off, rows, side=False):
        d = self._view(self.rank, dst[0], dst[1], dst_off, rows)
        sv = self._view(peer, src[0], src[1], src_off, rows)
        self._lane(side)
        try:
            d.copy_from(sv, rows=rows)
        finally:
            self._lane(False)

    def bcast_head(self, tile, rows):
        """Every GPU > 0 copies GPU 0's head tile (peer read over NVLink)."""
        if self.rank == 0:
            return
        d = self._view(self.rank, tile[0], tile[1], 0, rows)
        s = self._view(0, tile[0], tile[1], 0, rows)
        d.copy_from(s, rows=rows)

    def reduce_head(self, tile, rows):
        """GPU 0 pulls the partial head tiles of its peers and adds them (rank order => deterministic)."""
        if self.rank != 0:
            return
        d = self._view(0, tile[0], tile[1], 0, rows)
        ident = self._ident.get(rows)
        if ident is None:
            ident = self.ctx.map_upload(np.arange(rows, dtype=np.int64), rows)
            self._ident[rows] = ident
        for g in range(1, self.world):
            self.ctx.gather_rows(d, self._view(g, tile[0], tile[1], 0, rows), ident, accumulate=True)


class _CudaArray:
    """``__cuda_array_interface__`` view of a library-owned tile so torch.distributed can address it."""

    def __init__(self, ptr: int, rows: int, k: int):
        self.__cuda_array_interface__ = {"shape": (rows, k), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


class NcclBackend(CudaPeerBackend):
    """Same engine, but every cross-GPU step is an NCCL collective issued through torch.distributed:
    level exchange = pack kernel -> ``all_to_all_single`` -> unpack kernel, head tiles = ``broadcast`` / ``reduce``.
    This is the B200 version of the reference's own scheme (Alltoallv / Bcast / Reduce on host buffers,
    arrow_dec_mpi.py:442-505, arrow_slim_mpi.py:116, 273) and the A/B baseline for the peer-pull backend."""

    def __init__(self, comm, device: int, width: int, plan: ShardPlan):
        import torch
        self.torch = torch
        # run on torch's current stream so kernels and NCCL collectives are ordered; torch's default stream is the
        # legacy default stream (handle 0), which CUDA also names cudaStreamLegacy = 0x1
        super().__init__(comm, device, width, stream=torch.cuda.current_stream().cuda_stream or 1)
        self.plan_nccl = plan
        self.layout_plan = plan
        if not plan.block_diagonal:
            raise NotImplementedError("the NCCL backend covers the block-diagonal layout; use exchange='p2p' for banded")
        self._tables = {}
        self._bufs = {}

    def alloc_shared_tiles(self, rows_per_level, k, tiles_per_level=None):
        self.k = k
        tiles_per_level = list(tiles_per_level) if tiles_per_level is not None else [2] * len(rows_per_level)
        self._tiles = [[self.ctx.dense_alloc(r, k) for _ in range(nt)] for r, nt in zip(rows_per_level, tiles_per_level)]
        self._peer = [None] * self.world
        self._peer[self.rank] = self._tiles
        self._views = {}
        return self._tiles

    def barrier(self, side: bool = False):
        pass                                            # collectives carry the ordering

    def side_begin(self):
        pass

    def side_join(self):
        pass

    def _tensor(self, dense, rows):
        return self.torch.as_tensor(_CudaArray(dense.device_ptr(), rows, self.k), device="cuda")

    def _table(self, dst_level, forward):
        key = (dst_level, forward)
        t = self._tables.get(key)
        if t is None:
            raw = self.plan_nccl.a2a_tables(dst_level, forward)
            n_send, n_recv = int(raw["send_counts"].sum()), int(raw["recv_counts"].sum())
            src_level = dst_level - 1 if forward else dst_level + 1
            t = dict(send_counts=[int(c) for c in raw["send_counts"]], recv_counts=[int(c) for c in raw["recv_counts"]],
                     n_send=n_send, n_recv=n_recv,
                     pack=self.ctx.map_upload(raw["pack"], max(self.plan_nccl.levels[src_level].own_rows, 1)),
                     unpack=self.ctx.map_upload(raw["unpack"], max(n_recv, 1)),
                     sendbuf=self.ctx.dense_alloc(max(n_send, 1), self.k), recvbuf=self.ctx.dense_alloc(max(n_recv, 1), self.k))
            self._tables[key] = t
        return t

    def pull_rows(self, dst, dst_off, src, src_bounds, row_map, accumulate, forward=True, side=False):
        import torch.distributed as dist
        t = self._table(dst[0], forward)
        src_sh = self.plan_nccl.levels[src[0]]
        if t["n_send"] > 0:
            s_own = self._view(self.rank, src[0], src[1], src_sh.hoff, src_sh.own_rows)
            self.ctx.gather_rows(self._sub(t["sendbuf"], t["n_send"]), s_own, t["pack"])          # pack
        send_t = self._tensor(t["sendbuf"], t["n_send"])
        recv_t = self._tensor(t["recvbuf"], t["n_recv"])
        dist.all_to_all_single(recv_t, send_t, output_split_sizes=t["recv_counts"], input_split_sizes=t["send_counts"])
        n = row_map.n
        if n > 0 and t["n_recv"] > 0:
            d = self._view(self.rank, dst[0], dst[1], dst_off, n)
            self.ctx.gather_rows(d, self._sub(t["recvbuf"], t["n_recv"]), t["unpack"], accumulate=accumulate)   # unpack

    def _sub(self, dense, rows):
        key = ("sub", dense.h, rows)
        v = self._views.get(key)
        if v is None:
            v = self.ctx.dense_wrap(dense.device_ptr(), rows, self.k)
            self._views[key] = v
        return v

    def bcast_head(self, tile, rows):
        import torch.distributed as dist
        dist.broadcast(self._tensor(self._tiles[tile[0]][tile[1]], rows), src=0)

    def reduce_head(self, tile, rows):
        import torch.distributed as dist
        dist.reduce(self._tensor(self._tiles[tile[0]][tile[1]], rows), dst=0, op=dist.ReduceOp.SUM)


class ShardedArrowDecomposition:
    """Convenience wrapper used by bench.py at N > 1: plan + CUDA peer backend + the reference-like calls."""

    def __init__(self, comm, decomposition, width: int, k: int, device: int = 0, exchange: str = "p2p",
                 overlap: bool = False, block_diagonal: bool = True, mode: str = "auto"):
        self.comm = comm
        plan = ShardPlan(decomposition, width, comm.Get_rank(), comm.Get_size(), block_diagonal=block_diagonal)
        if exchange == "p2p":
            be = CudaPeerBackend(comm, device, width, plan=plan)
        elif exchange == "p2p-direct":
            be = CudaPeerBackend(comm, device, width)
            be.layout_plan = plan
        elif exchange == "nccl":
            be = NcclBackend(comm, device, width, plan)
        else:
            raise ValueError("exchange must be 'p2p', 'p2p-direct' or 'nccl'")
        self.engine = ShardedArrowEngine(plan, k, be, overlap=overlap, mode=mode)
        self.B = self
        self.matrix_index = 0
        self.decomposition_length = plan.L

    def set_features(self, X):
        self.engine.set_features(np.ascontiguousarray(X, dtype=np.float32))

    def result_tile(self, out=None):
        return self.engine.result(0, out)

    def step(self):
        self.engine.step()

    def step_stream(self, X_host, out_host):
        self.engine.stream_step(X_host, out_host)

    def synchronize(self):
        self.engine.sync()
