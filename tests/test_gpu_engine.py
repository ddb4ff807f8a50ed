"""GPU parity of the whole iteration (ArrowEngine.step) against the protocol oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle
from arrow_matrix_b200 import _lib, synth
from arrow_matrix_b200.engine import ArrowEngine
from tests.test_gpu_kernels import assert_close


@pytest.mark.parametrize("mode", ["fused", "exchange"])
@pytest.mark.parametrize("perm_kind", ["random", "local", "identity"])
@pytest.mark.parametrize("k,levels", [(16, 2), (128, 2), (4, 3), (10, 3)])
def test_step_matches_protocol_oracle_chained(cuda_device, mode, perm_kind, k, levels):
    w, t0 = 64, 12
    dec = synth.synth_decomposition(t0, w, levels=levels, perm_kind=perm_kind, seed=21, hub_rows=2, hub_nnz=700)
    n = t0 * w
    eng = ArrowEngine(dec, w, k, device=cuda_device, mode=mode)
    assert eng.mode == mode
    po = oracle.ReferenceProtocolOracle(dec, w, k)
    po64 = oracle.ReferenceProtocolOracle(dec, w, k, dtype=np.float64)      # exact yardstick (see assert_close)
    X = synth.generate_dense_matrix(n, k, np.float32, np.random.default_rng(42))
    Xl0 = X[po.perms[0]]
    eng.set_features(Xl0)
    po.set_features(Xl0.copy())
    po64.set_features(Xl0)
    for it in range(3):                                   # chained like tests/test_arrowmpi.py:164-166
        eng.step()
        ref = po.step()
        got = eng.result()
        assert_close(got, ref, exact=po64.step())
        # re-sync the oracles' state to the device result so errors do not compound across iterations
        po.C[0][:] = got
        po64.C[0][:] = got
    # against the reference tests' own golden for a fresh X
    eng.set_features(Xl0)
    eng.step()
    gold = oracle.compute_spmm(dec, X)
    assert_close(oracle.to_original_order(eng.result(), po.perms[0], n), gold)
    eng.close()


def test_exchange_mode_level_tiles_and_stale_rows(cuda_device):
    """non-nested permutations: rows behind the sentinel keep the previous result (arrow_dec_mpi.py:544)."""
    w, t0, k = 32, 8, 8
    dec = synth.synth_decomposition(t0, w, levels=3, perm_kind="random", seed=4, nested=False)
    eng = ArrowEngine(dec, w, k, device=cuda_device, mode="auto")
    assert eng.mode == "exchange" and not eng.fused_ok
    with pytest.raises(ValueError):
        ArrowEngine(dec, w, k, device=cuda_device, mode="fused")
    po = oracle.ReferenceProtocolOracle(dec, w, k)
    rng = np.random.default_rng(1)
    for it in range(3):
        X = synth.generate_dense_matrix(t0 * w, k, np.float32, rng)   # fresh X per iteration like arrow_bench.py:113-116
        eng.set_features(X)
        po.set_features(X.copy())
        eng.step()
        po.step()
        for j in range(3):
            assert_close(eng.result(j), po.C[j])
    # after step() + _propagate_features() every level holds its permuted slice (test_arrowmpi.py:283, 306-309)
    eng.propagate_features()
    po.propagate_features()
    for j in range(3):
        assert_close(eng.result(j), po.C[j])
    eng.close()


@pytest.mark.parametrize("levels,k", [(2, 32), (3, 16), (3, 6), (4, 128)])
def test_fused_styles_equal_exchange(cuda_device, levels, k):
    """fused/gather (epilogue gather-add), fused/scatter (row-map epilogue) and the literal exchange agree"""
    w, t0 = 100, 16
    dec = synth.synth_decomposition(t0, w, levels=levels, perm_kind="random", seed=8, hub_rows=2, hub_nnz=900)
    X = synth.generate_dense_matrix(t0 * w, k, np.float32, np.random.default_rng(0))
    res = {}
    for name, kw in (("gather", dict(mode="fused", fused_style="gather")), ("scatter", dict(mode="fused", fused_style="scatter")),
                     ("exchange", dict(mode="exchange"))):
        eng = ArrowEngine(dec, w, k, device=cuda_device, **kw)
        eng.set_features(X)
        eng.step()
        first = eng.result()
        eng.step()                  # chained second iteration
        res[name] = (first, eng.result())
        eng.close()
    # exact yardstick: two float64 steps, the second one started from each engine's own first result
    for name in ("gather", "scatter"):
        assert_close(res[name][0], res["exchange"][0])
        po64 = oracle.ReferenceProtocolOracle(dec, w, k, dtype=np.float64)
        po64.set_features(X)
        po64.step()
        po64.C[0][:] = res["exchange"][0]
        po64.X[0] = po64.C[0]
        assert_close(res[name][1], res["exchange"][1], exact=po64.step())


def test_arrow_pattern_masking(cuda_device):
    """non-zeros outside the arrow pattern are dropped exactly like graphio.split_matrix_to_blocks (:382-383)."""
    from scipy import sparse
    w, t0, k = 16, 5, 4
    dec = synth.synth_decomposition(t0, w, levels=1, seed=3)
    B = sparse.lil_matrix(dec[0][0])
    B[3 * w + 1, 1 * w + 2] = 5.0          # block (3,1): outside the pattern
    B[2 * w, 4 * w + 3] = 2.0              # block (2,4): outside
    B = sparse.csr_matrix(B)
    dec2 = [(B, dec[0][1])]
    eng = ArrowEngine(dec2, w, k, device=cuda_device)
    assert eng.levels[0].dropped == 2
    po = oracle.ReferenceProtocolOracle(dec2, w, k)
    assert po.dropped_nnz == [2]
    X = synth.generate_dense_matrix(t0 * w, k, np.float32, np.random.default_rng(5))
    eng.set_features(X)
    po.set_features(X)
    eng.step()
    assert_close(eng.result(), po.step())
    eng.close()


# ---- against golden vectors produced by the real reference (tests/golden/make_golden.py) ----------------
from tests.golden_util import GPU_CASES, GoldenCase


@pytest.mark.parametrize("name", GPU_CASES)
def test_engine_matches_real_reference_run(cuda_device, name):
    g = GoldenCase(name)
    eng = ArrowEngine(g.decomposition, g.width, g.k, block_diagonal=g.block_diagonal, device=cuda_device, mode="exchange")
    assert eng.n_blocks == g.n_blocks
    for it in range(g.iterations):
        if g.X[it] is not None:
            eng.set_features(g.X[it])
        eng.step()
        for j in range(g.L):
            assert_close(eng.result(j), g.C[it][j])
    eng.propagate_features()
    for j in range(g.L):
        assert_close(eng.result(j), g.final[j])
    fused_ok = eng.fused_ok
    eng.close()
    if fused_ok:
        eng = ArrowEngine(g.decomposition, g.width, g.k, block_diagonal=g.block_diagonal, device=cuda_device, mode="fused")
        for it in range(g.iterations):
            if g.X[it] is not None:
                eng.set_features(g.X[it])
            eng.step()
            assert_close(eng.result(0), g.C[it][0])
        eng.close()


@pytest.mark.parametrize("mode", ["fused", "exchange"])
def test_stream_step_matches_blocking_calls(cuda_device, mode):
    """pipelined host-staged iterations (copy lanes + events) give the same tiles as set_features/step/result"""
    w, t0, k = 64, 10, 32
    dec = synth.synth_decomposition(t0, w, levels=2, perm_kind="random", seed=12)
    n = t0 * w
    rng = np.random.default_rng(3)
    Xs = [synth.generate_dense_matrix(n, k, np.float32, rng) for _ in range(5)]
    eng = ArrowEngine(dec, w, k, device=cuda_device, mode=mode)
    ref = []
    for X in Xs:
        eng.set_features(X)
        eng.step()
        ref.append(eng.result())
    eng.close()
    eng = ArrowEngine(dec, w, k, device=cuda_device, mode=mode)
    hx = [_lib.PinnedArray((n, k)) for _ in range(2)]
    hc = [_lib.PinnedArray((n, k)) for _ in range(2)]
    got = []
    for i, X in enumerate(Xs):
        if i >= 2:
            eng.stream_drain()                  # host buffers of call i-2 are about to be reused
            got.append(hc[i % 2].array.copy())
        hx[i % 2].array[:] = X
        eng.stream_step(hx[i % 2].array, hc[i % 2].array)
    eng.stream_drain()
    # collect the last two (drained above only up to i-2)
    got = got[:len(Xs) - 2] + [hc[(len(Xs) - 2) % 2].array.copy(), hc[(len(Xs) - 1) % 2].array.copy()]
    for g, r in zip(got, ref):
        assert np.array_equal(g, r)
    eng.close()


@pytest.mark.parametrize("n,w,k", [(1000, 64, 8), (630, 100, 16)])
def test_decomposed_graph_through_engine(cuda_device, n, w, k):
    """graph -> igraph-free arrow decomposition -> device engine == the reference protocol on the same levels
    (ragged last block, non-nested permutations, best-effort last level)"""
    from arrow_matrix_b200.decomposition import arrow_decomposition
    A = synth.barabasi_albert(n, 4, seed=9)
    dec = arrow_decomposition(A, w, max_number_of_levels=3, block_diagonal=True, seed=1)
    eng = ArrowEngine(dec, w, k, device=cuda_device, mode="auto")
    po = oracle.ReferenceProtocolOracle(dec, w, k)
    po64 = oracle.ReferenceProtocolOracle(dec, w, k, dtype=np.float64)      # exact yardstick (hub rows of a BA graph)
    rng = np.random.default_rng(6)
    X = synth.generate_dense_matrix(po.rows[0], k, np.float32, rng)
    eng.set_features(X)
    po.set_features(X.copy())
    po64.set_features(X)
    for it in range(2):
        eng.step()
        po.step()
        po64.step()
        for j in range(po.L if eng.mode == "exchange" else 1):      # fused mode keeps only level 0's tile
            assert_close(eng.result(j), po.C[j], exact=po64.C[j])
        po64.C[0][:] = po.C[0]              # the second (chained) product is compared from the same starting point
    eng.close()


def test_ones_step_property_at_benchmark_scale(cuda_device):
    """the size-independent parity property bench.py attaches to every run ("verified"): one step on all-ones features
    equals the row sums of every level pushed through the exchange maps.  1M rows by default, the BASELINE.json size
    (10M rows, width 10 000, k = 128) with ARROW_TEST_FULL_SIZE=1."""
    import os
    import bench
    from arrow_matrix_b200.comm import SelfComm
    full = os.environ.get("ARROW_TEST_FULL_SIZE") == "1"
    blocks, w, k = (1000, 10000, 128) if full else (100, 10000, 128)
    dec = synth.synth_decomposition(blocks, w, levels=2, perm_kind="random", seed=503)
    eng = ArrowEngine(dec, w, k, device=cuda_device)
    hx, hc = _lib.PinnedArray((blocks * w, k)), _lib.PinnedArray((blocks * w, k))
    v = bench.verify_ones_step(eng, dec, w, 0, hx, hc, SelfComm())
    v1 = bench.verify_rank1_step(eng, dec, w, 0, hx, hc, SelfComm())        # the property the bench line carries since round 2
    eng.close()
    assert v.get("ok") is True, v
    assert v["rows"] == blocks * w and v["max_rel_err"] <= 1e-5
    assert v1.get("ok") is True and v1["rows"] == blocks * w and v1["max_rel_err"] <= 1e-5, v1


# ---- BASELINE.json configurations with real values (round 2) --------------------------------------
def test_baseline_config1_100k_rows_k16_two_levels_chained(cuda_device, tmp_path):
    """BASELINE.json configs[0]: random 100k-row, width 10 000, k = 16 decomposition, through the level files and the public
    classes, three chained iterations like the reference's own test (tests/test_arrowmpi.py:164-166), against the oracle"""
    from arrow_matrix_b200 import graphio
    from arrow_matrix_b200.arrow_dec_mpi import ArrowDecompositionMPI
    from arrow_matrix_b200.comm import SelfComm
    w, t0, k = 10000, 10, 16
    dec = synth.synth_decomposition(t0, w, levels=2, perm_kind="random", seed=503)
    base = str(tmp_path / "cfg1")
    graphio.save_decomposition_new(dec, base, w, block_diagonal=True)
    comm = SelfComm()
    blocks, n_blocks, to_prev, to_next = ArrowDecompositionMPI.load_decomposition_new(comm, base, w, True, slim=True)
    arrow = ArrowDecompositionMPI.initialize(comm, n_blocks, to_prev, to_next, w, k, 'gpu', True, True)
    arrow.B.load_sparse_matrix_from_blocks(blocks)
    arrow.B.zero_rhs(w, k)
    po = oracle.ReferenceProtocolOracle(dec, w, k, use_c_kernel=True)
    po64 = oracle.ReferenceProtocolOracle(dec, w, k, dtype=np.float64)
    X = synth.generate_dense_matrix(t0 * w, k, np.float32, np.random.default_rng(42))
    arrow.B.set_features(X)
    po.set_features(X.copy())
    po64.set_features(X)
    for it in range(3):
        arrow.step()
        ref = po.step()
        got = arrow.B.result_tile()
        assert_close(got, ref, exact=po64.step())
        po.C[0][:] = got                                   # compare every product from the same starting point
        po64.C[0][:] = got
    arrow._engine.close()


@pytest.mark.parametrize("recipe", ["uniform", "arrow"])
def test_baseline_configs_2_3_one_million_row_block_k_sweep(cuda_device, recipe):
    """BASELINE.json configs[1] and [2]: one 1M x 1M CSR with ~10 non-zeros per row times dense k in {16, 32, 64, 128, 256},
    fp32 -- the reference's generator recipe (uniform columns, arrow/common/utils.py:63-87, rng 42) and the arrow-structured
    variant of SURVEY 8d -- against the restated SciPy kernel (oracle/csr_matvecs.c), 1e-5 of the largest entry"""
    n = 1_000_000
    rng = np.random.default_rng(42)
    A = synth.generate_sparse_matrix(n, n, 10 * n, np.float32, rng) if recipe == "uniform" else synth.arrow_csr(n, 10000, 100, rng)
    ctx = _lib.Context(cuda_device)
    dA = ctx.csr_from_scipy(A)
    for k in (16, 32, 64, 128, 256):
        X = synth.generate_dense_matrix(n, k, np.float32, rng)
        dX, dC = ctx.dense_from_host(X), ctx.dense_alloc(n, k)
        ctx.spmm(dA, dX, dC)
        got = dC.d2h()
        ref = oracle.csr_spmm_c(A, X)
        scale = float(np.max(np.abs(ref)))
        err = float(np.max(np.abs(got - ref)))
        assert err <= 1e-5 * scale, (recipe, k, err / scale)
        dX.free()
        dC.free()
    ctx.close()
