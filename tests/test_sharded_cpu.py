"""CPU tests of the multi-GPU path's host logic: world_size 2 and 3 over gloo, numpy tiles.

Each rank builds its ShardPlan, runs ShardedArrowEngine.step() on the gloo/numpy backend and compares
its own rows with the protocol oracle (which itself is pinned against the real reference)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cases, q):
    """one process per rank; the cases of one world size share the process group (a spawn + torch import per case
    would dominate the suite's run time).  A failing rank reports and exits: its peers then fail fast on the closed
    connection instead of waiting for it."""
    current = None
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        for case, overlap in cases:
            current = (case, overlap)
            _run_case(rank, world, case, overlap)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException:     # noqa: BLE001
        import traceback
        q.put((rank, f"FAIL in case {current}: " + traceback.format_exc()))


def _run_case(rank, world, case, overlap):
    from arrow_matrix_b200 import synth
    from arrow_matrix_b200.comm import TorchComm, world_comm
    from arrow_matrix_b200.sharded import ShardPlan, ShardedArrowEngine
    from oracle import oracle
    from tests.numpy_backend import GlooNumpyBackend
    comm = world_comm()
    assert isinstance(comm, TorchComm) and comm.Get_size() == world and comm.Get_rank() == rank
    if case.startswith("golden:"):
        from tests.golden_util import GoldenCase
        g = GoldenCase(case.split(":", 1)[1])
        dec, w, k = g.decomposition, g.width, g.k
        Xs = g.X
        bd = g.block_diagonal
    elif case.startswith("decomposed"):
        # a graph run through the igraph-free arrow decomposition (n is not a multiple of the width: ragged tail)
        from arrow_matrix_b200.decomposition import arrow_decomposition
        n, w, k = (630, 100, 8) if case == "decomposed" else (1000, 64, 4)
        A = synth.barabasi_albert(n, 4, seed=9)
        dec = arrow_decomposition(A, w, max_number_of_levels=3, block_diagonal=True, seed=1)
        rows0 = oracle.number_of_blocks(dec[0][0], w) * w
        rng = np.random.default_rng(6)
        Xs = [synth.generate_dense_matrix(rows0, k, np.float32, rng), None]
    else:
        w, t0, k, levels, kind, nested = {"L2": (16, 7, 8, 2, "random", True), "L3": (8, 9, 5, 3, "random", True),
                                          "L3stale": (8, 6, 4, 3, "random", False), "small": (8, 2, 4, 2, "random", True),
                                          "banded": (8, 9, 4, 2, "random", True), "local16": (8, 16, 4, 2, "local", True)}[case]
        dec = synth.synth_decomposition(t0, w, levels=levels, perm_kind=kind, seed=77, nested=nested, hub_rows=2, hub_nnz=40,
                                        band_nnz=3 if case == "banded" else 0, shrink=1 if case == "banded" else 2)
        rng = np.random.default_rng(5)
        n0 = t0 * w
        Xs = [synth.generate_dense_matrix(n0, k, np.float32, rng), None, synth.generate_dense_matrix(n0, k, np.float32, rng)]
    bd = locals().get("bd", True)
    if case == "banded":
        bd = False
    plan = ShardPlan(dec, w, rank, world, block_diagonal=bd)
    if case == "local16":
        # shard-local permutation: level 1 is cut where its rows map, so (apart from the head rows) nothing is exchanged
        assert plan.partition_used == ["even", "locality"], plan.partition_used
        from arrow_matrix_b200.sharded import FusedPlan
        assert FusedPlan(plan).recv_rows <= w
        even = comm.allgather(FusedPlan(ShardPlan(dec, w, rank, world, block_diagonal=bd, partition="even")).recv_rows)
        assert max(even) > w, even                          # the even split would have sent whole shards across
    fused = isinstance(overlap, str)            # "fused" / "fused-side": the fused step (serial / side-lane schedule)
    if fused:
        eng = ShardedArrowEngine(plan, k, GlooNumpyBackend(comm, w, plan), overlap=(overlap == "fused-side"), mode="auto")
        if eng.fp is not None and overlap == "fused-side":
            eng.bwd_mode = "pull"               # both backward transports: SM push (default) and copy-engine pulls
        # rows behind the sentinel only force the literal protocol when some non-zero READS one of them (then the
        # engine must have fallen back on its own); either way the numbers below have to match the oracle
        if case == "L3stale":
            assert not eng.fused_ok and eng.fp is None, (case, eng.mode)
        elif not case.endswith(("nonnested_k3", "stale_k7")):
            assert eng.fused_ok and eng.fp is not None, (case, eng.mode)
    else:
        eng = ShardedArrowEngine(plan, k, GlooNumpyBackend(comm, w, plan), overlap=overlap, mode="exchange")
        assert eng.overlap == bool(overlap)
        assert eng.split == (overlap == 2 and plan.L == 2 and world > 1)
    po = oracle.ReferenceProtocolOracle(dec, w, k, block_diagonal=bd)
    assert eng.total_nnz == sum(M.nnz for M in po.mats)
    sh0 = plan.levels[0]
    for it, X in enumerate(Xs):
        if X is not None:
            eng.set_features(X[sh0.r0:sh0.r1])
            po.set_features(X.copy())
        eng.step()
        po.step()
        for j in range(plan.L if eng.fp is None else 1):        # the fused step materialises level 0 only
            sh = plan.levels[j]
            got = eng.result(j)
            assert got.shape == (sh.own_rows, k)
            assert np.allclose(got, po.C[j][sh.r0:sh.r1], rtol=1e-5, atol=1e-5), (case, rank, it, j)
    if eng.fp is not None:
        with pytest.raises(RuntimeError):
            eng.result(1)
        # the phase-by-phase surface falls back to the literal protocol and keeps producing the same numbers
        eng.ensure_level_tiles()
        assert eng.fp is None
        eng.step()
        po.step()
        for j in range(plan.L):
            sh = plan.levels[j]
            ref = po.C[j][sh.r0:sh.r1]
            err = float(np.max(np.abs(eng.result(j) - ref))) / max(float(np.max(np.abs(po.C[j]))), 1e-30) if ref.size else 0.0
            assert err <= 1e-5, (case, rank, "after fallback", j, err)      # third chained product: max-norm relative
    assert comm.allreduce_lor(False) is False


CASES = [(w, c, False) for w in (2, 3) for c in ["L2", "L3", "L3stale", "small", "golden:slim_L2_random_k4",
                                                 "golden:slim_L3_nonnested_k3"]] + \
        [(2, "L3", True), (3, "L2", True), (3, "L3stale", True)] + \
        [(2, "banded", False), (3, "banded", True), (4, "banded", False),
         (2, "golden:wide_L2_banded_k4", False), (3, "golden:wide_L2_banded_k4", True),
         (2, "decomposed", False), (3, "decomposed-1000", True),
         (3, "golden:slim_L4_nested_k6", True), (2, "golden:wide_L3_banded_stale_k7", False),
         (3, "golden:wide_L3_banded_stale_k7", True), (4, "golden:slim_L4_nested_k6", False),
         # overlap=2: split level-0 product, staged backward exchange (two levels; else falls back)
         (2, "L2", 2), (3, "L2", 2), (4, "small", 2), (3, "golden:slim_L2_random_k4", 2),
         (3, "golden:slim_L2_short_file_k4", 2), (3, "L3", 2),
         # the fused step (push exchange folded into the products), serial and side-lane schedules; stale-row
         # decompositions must fall back to the literal protocol on their own
         (2, "L2", "fused"), (2, "L3", "fused-side"), (2, "small", "fused"), (2, "banded", "fused"), (2, "L3stale", "fused"),
         (2, "golden:slim_L2_hubs_k16", "fused-side"), (2, "golden:slim_L4_nested_k6", "fused"),
         (2, "golden:wide_L2_banded_k4", "fused"), (2, "decomposed", "fused-side"),
         (3, "L2", "fused-side"), (3, "L3", "fused"), (3, "banded", "fused-side"), (3, "golden:slim_L2_random_k4", "fused"),
         (3, "golden:slim_L3_nonnested_k3", "fused"), (3, "golden:slim_L2_julia_quirks_k4", "fused"),
         (3, "decomposed-1000", "fused"), (3, "golden:wide_L3_banded_stale_k7", "fused-side"),
         (4, "L2", "fused"), (4, "small", "fused-side"), (4, "banded", "fused"), (4, "golden:slim_L4_nested_k6", "fused-side"),
         (4, "golden:wide_L2_random_k5", "fused"),
         (2, "local16", "fused"), (2, "local16", False), (4, "local16", "fused-side"), (4, "local16", True)]


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_engine_over_gloo(world):
    import torch.multiprocessing as mp
    cases = [(c, o) for w, c, o in CASES if w == world]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(30)
    bad = [f"rank {rank}: {msg}" for rank, msg in sorted(results) if msg != "ok" and "Connection closed by peer" not in msg]
    bad = bad or [f"rank {rank}: {msg}" for rank, msg in sorted(results) if msg != "ok"]
    assert not bad, "\n".join(bad)


def test_shard_plan_covers_the_matrix():
    """the local matrices of all ranks together hold every arrow entry exactly once"""
    sys.path.insert(0, ROOT)
    from arrow_matrix_b200 import synth
    from arrow_matrix_b200.sharded import ShardPlan
    w, t0 = 8, 7
    dec = synth.synth_decomposition(t0, w, levels=2, seed=3, hub_rows=2, hub_nnz=30)
    for world in (1, 2, 4, 8, 16):
        tot = [0, 0]
        for r in range(world):
            pl = ShardPlan(dec, w, r, world)
            for j in range(2):
                tot[j] += pl.levels[j].nnz
                assert pl.levels[j].indices.size == pl.levels[j].nnz
                if pl.levels[j].nnz:
                    assert pl.levels[j].indices.max() < pl.levels[j].local_rows
        assert tot == [dec[0][0].nnz, dec[1][0].nnz], (world, tot)


def test_a2a_tables_simulated_exchange():
    """pack -> all-to-all -> unpack with the plan's tables delivers exactly dst[r] = src[map[r]] (NCCL / packed backends)"""
    sys.path.insert(0, ROOT)
    from arrow_matrix_b200 import synth
    from arrow_matrix_b200.sharded import ShardPlan
    w, t0 = 8, 9
    dec = synth.synth_decomposition(t0, w, levels=3, perm_kind="random", seed=5, nested=False)
    for world in (1, 2, 3, 4):
        plans = [ShardPlan(dec, w, r, world) for r in range(world)]
        for (lvl, fwd) in [(1, True), (2, True), (0, False), (1, False)]:
            T = [p.a2a_tables(lvl, fwd) for p in plans]
            src_level = lvl - 1 if fwd else lvl + 1
            src_tiles = [np.arange(p.levels[src_level].r0, p.levels[src_level].r1) for p in plans]   # rows hold their global id
            send = [src_tiles[r][T[r]["pack"]] for r in range(world)]
            recv = [[] for _ in range(world)]
            for s in range(world):
                off = 0
                for d in range(world):
                    c = int(T[s]["send_counts"][d])
                    recv[d].append((s, send[s][off:off + c]))
                    off += c
            for d in range(world):
                parts = sorted(recv[d], key=lambda x: x[0])
                assert [len(x[1]) for x in parts] == [int(c) for c in T[d]["recv_counts"]]
                rb = np.concatenate([x[1] for x in parts])
                p = plans[d]
                sh = p.levels[lvl]
                gm = (p.to_prev[lvl] if fwd else p.to_next[lvl])[sh.r0:sh.r1]
                exp = np.where(gm < p.levels[src_level].rows_global, gm, -1)
                got = np.where(T[d]["unpack"] >= 0, rb[np.maximum(T[d]["unpack"], 0)] if rb.size else -1, -1)
                assert np.array_equal(got, exp), (world, lvl, fwd, d)
                # receivers read every peer region front to back: positions ascend with the destination row
                for s in range(world):
                    lo = int(np.sum(T[d]["recv_counts"][:s]))
                    hi = lo + int(T[d]["recv_counts"][s])
                    pos = T[d]["unpack"][(T[d]["unpack"] >= lo) & (T[d]["unpack"] < hi)]
                    assert np.all(np.diff(pos) > 0)


@pytest.mark.parametrize("block_diagonal", [True, False])
def test_shard_plan_algebra_without_comm(block_diagonal):
    """local matrices x local tiles (head tile, halos, own rows) re-assemble the level's arrow product exactly"""
    sys.path.insert(0, ROOT)
    from scipy import sparse
    from arrow_matrix_b200 import synth
    from arrow_matrix_b200.sharded import ShardPlan
    from oracle import oracle
    w, t0, k = 8, 9, 3
    dec = synth.synth_decomposition(t0, w, levels=2, seed=13, hub_rows=2, hub_nnz=30,
                                    band_nnz=0 if block_diagonal else 3, shrink=1)
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 5, 9, 12):
        plans = [ShardPlan(dec, w, r, world, block_diagonal=block_diagonal) for r in range(world)]
        for j in range(2):
            nb = plans[0].n_blocks[j]
            M = oracle.arrow_mask(dec[j][0], w, nb, block_diagonal)
            X = rng.random((nb * w, k), dtype=np.float32).astype(np.float64)
            ref = M.astype(np.float64) @ X
            C = np.zeros_like(ref)
            head = np.zeros((w, k))
            for r, pl in enumerate(plans):
                sh = pl.levels[j]
                Xl = np.zeros((sh.local_rows, k))
                if r > 0:
                    Xl[:w] = X[:w]
                Xl[sh.hoff:sh.hoff + sh.own_rows] = X[sh.r0:sh.r1]
                if sh.halo_prev_off >= 0:
                    g, first = sh.halo_prev_src
                    assert plans[g].levels[j].r0 <= first < plans[g].levels[j].r1
                    Xl[sh.halo_prev_off:sh.halo_prev_off + w] = X[first:first + w]
                if sh.halo_next_off >= 0:
                    g, first = sh.halo_next_src
                    assert plans[g].levels[j].r0 <= first < plans[g].levels[j].r1
                    Xl[sh.halo_next_off:sh.halo_next_off + w] = X[first:first + w]
                Ml = sparse.csr_matrix((sh.data.astype(np.float64), sh.indices, sh.indptr), shape=(sh.local_rows, sh.local_rows))
                Cl = Ml @ Xl
                C[sh.r0:sh.r1] = Cl[sh.hoff:sh.hoff + sh.own_rows]
                if r > 0:
                    head += Cl[:w]                       # partial C_0 of this rank (Reduce to rank 0)
                assert pl.hoff_of(j, r) == sh.hoff
            C[:w] += head
            assert np.allclose(C, ref, rtol=1e-12, atol=1e-12), (world, j)


def test_shard_plan_from_memory_mapped_files(tmp_path):
    """the public multi-GPU path hands ShardPlan the memory-mapped npy triplets: same shards as from scipy matrices,
    also without a data file (ones) and with int64 indices"""
    sys.path.insert(0, ROOT)
    from arrow_matrix_b200 import graphio, synth
    from arrow_matrix_b200.sharded import ShardPlan
    w, t0 = 8, 6
    dec = synth.synth_decomposition(t0, w, levels=2, seed=21, hub_rows=2, hub_nnz=20)
    base = str(tmp_path / "g")
    graphio.save_decomposition_new(dec, base, w, True)
    mm = graphio.load_decomposition_new(base, w, True, mem_map=True)
    base2 = str(tmp_path / "jl")
    graphio.save_decomposition_new(dec, base2, w, True, write_data=False, index_dtype=np.int64)
    mm2 = graphio.load_decomposition_new(base2, w, True, mem_map=True)
    for world in (1, 3):
        for r in range(world):
            a = ShardPlan(dec, w, r, world)
            b = ShardPlan(mm, w, r, world)
            c = ShardPlan(mm2, w, r, world)
            for j in range(2):
                sa, sb, sc = a.levels[j], b.levels[j], c.levels[j]
                assert np.array_equal(sa.indptr, sb.indptr) and np.array_equal(sa.indices, sb.indices)
                assert np.array_equal(sa.data, sb.data)
                assert np.array_equal(sa.indptr, sc.indptr) and np.array_equal(sa.indices, sc.indices)
                assert np.all(sc.data == 1.0) and sc.data.dtype == np.float32
                assert np.array_equal(sa.fwd_map if sa.fwd_map is not None else [], sb.fwd_map if sb.fwd_map is not None else [])


@pytest.mark.parametrize("block_diagonal,band", [(True, 0), (False, 3)])
def test_shard_local_matrices_are_valid_uploads(block_diagonal, band):
    """what every rank hands to arrow_csr_upload: a row pointer spanning exactly its entries and columns inside
    [0, local_rows) -- the C ABI rejects anything else (also for ranks that own nothing)"""
    sys.path.insert(0, ROOT)
    from arrow_matrix_b200 import synth
    from arrow_matrix_b200.sharded import ShardPlan
    for (t0, w, levels, shrink) in ((7, 8, 2, 2), (9, 8, 3, 1), (2, 8, 2, 1), (5, 4, 3, 1)):
        dec = synth.synth_decomposition(t0, w, levels=levels, seed=3, hub_rows=2, hub_nnz=20, band_nnz=band, shrink=shrink)
        for world in (1, 2, 3, 8, 16):
            for r in range(world):
                pl = ShardPlan(dec, w, r, world, block_diagonal=block_diagonal)
                for sh in pl.levels:
                    assert sh.indptr.size == sh.local_rows + 1 and int(sh.indptr[-1]) == sh.indices.size == sh.nnz
                    assert np.all(np.diff(sh.indptr) >= 0)
                    if sh.nnz:
                        assert sh.indices.min() >= 0 and sh.indices.max() < sh.local_rows


def test_locality_partition_rules():
    from arrow_matrix_b200 import decomp
    w, nb, parts = 4, 8, 4
    prev = np.array([0, 16, 32, 48, 64])                       # level above: 16 rows per GPU
    ident = np.arange(nb * w)
    # level rows map onto the first half of the level above, in order: blocks 0-3 -> GPU 0, 4-7 -> GPU 1
    assert decomp.locality_partition(ident, nb, w, prev, parts).tolist() == [0, 4, 8, 8, 8]
    # a uniformly random permutation keeps the even split
    assert decomp.locality_partition(np.random.default_rng(0).permutation(64)[:32], nb, w, prev, parts) is None
    # votes that are not monotone cannot give contiguous shards
    rev = ident[::-1].copy()
    assert decomp.locality_partition(rev + 32, nb, w, prev, parts) is None
    # block-row 0 stays on GPU 0 even when its rows map elsewhere
    shifted = np.concatenate([np.arange(48, 52), np.arange(4, 32)])
    b = decomp.locality_partition(shifted, nb, w, prev, parts)
    assert b is not None and b[1] >= 1
    # one GPU would get everything: too skewed
    assert decomp.locality_partition(ident % 16, nb, w, prev, parts, max_skew=2.0) is None


@pytest.mark.parametrize("world,case", [(8, "L2"), (8, "L3"), (8, "banded"), (5, "L2"), (7, "L3"), (8, "local16"), (6, "golden:slim_L4_nested_k6")])
def test_fused_step_routing_up_to_eight_ranks(world, case):
    """the fused step's routing (rotated destination lists, staging / send tile layout, head-row delivery) for the world sizes the
    benchmark runs at: rank threads in this process over the numpy test double (remote stores are delivered at barriers, remote reads
    come from snapshots), both backward transports, every rank against the protocol oracle"""
    import threading
    sys.path.insert(0, ROOT)
    from arrow_matrix_b200 import synth
    from arrow_matrix_b200.comm import ThreadWorld
    from arrow_matrix_b200.sharded import ShardPlan, ShardedArrowEngine
    from oracle import oracle
    from tests.numpy_backend import GlooNumpyBackend
    if case.startswith("golden:"):
        from tests.golden_util import GoldenCase
        g = GoldenCase(case.split(":", 1)[1])
        dec, w, k, bd = g.decomposition, g.width, g.k, g.block_diagonal
    else:
        w, t0, k, levels, kind = {"L2": (8, 19, 4, 2, "random"), "L3": (8, 21, 3, 3, "random"), "banded": (8, 17, 4, 2, "random"),
                                  "local16": (8, 16, 4, 2, "local")}[case]
        bd = case != "banded"
        dec = synth.synth_decomposition(t0, w, levels=levels, perm_kind=kind, seed=77, hub_rows=2, hub_nnz=40,
                                        band_nnz=0 if bd else 3, shrink=2 if bd else 1)
    po = oracle.ReferenceProtocolOracle(dec, w, k, block_diagonal=bd)
    rng = np.random.default_rng(5)
    Xs = [synth.generate_dense_matrix(po.rows[0], k, np.float32, rng) for _ in range(2)]
    refs = []
    for X in Xs:
        po.set_features(X.copy())
        refs.append(po.step().copy())
        refs.append(po.step().copy())                       # a chained step after every fresh one
    tw = ThreadWorld(world)
    errors = [None] * world

    def body(rank):
        comm = tw.comm(rank)
        try:
            plan = ShardPlan(dec, w, rank, world, block_diagonal=bd)
            eng = ShardedArrowEngine(plan, k, GlooNumpyBackend(comm, w, plan), overlap=bool(world % 3), mode="fused")
            eng.bwd_mode = "pull" if world % 2 else "push"
            sh0 = plan.levels[0]
            i = 0
            for X in Xs:
                eng.set_features(X[sh0.r0:sh0.r1])
                for _ in range(2):
                    eng.step()
                    got = eng.result(0)
                    assert np.allclose(got, refs[i][sh0.r0:sh0.r1], rtol=1e-5, atol=1e-5 * max(1.0, float(np.max(np.abs(refs[i]))))), (rank, i)
                    i += 1
        except BaseException:     # noqa: BLE001
            import traceback
            errors[rank] = traceback.format_exc()
            comm.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    real = [e for e in errors if e and "BrokenBarrierError" not in e]
    assert not real and not any(errors), "\n".join(real or [e for e in errors if e])
