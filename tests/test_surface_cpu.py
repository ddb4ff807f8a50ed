"""CPU tests of the host-side surface: C-ABI exports, loud failure without a GPU, loader, routing tables."""
import os
import re

import numpy as np
import pytest

from arrow_matrix_b200 import _lib, graphio, synth
from arrow_matrix_b200.arrow_dec_mpi import ArrowDecompositionMPI
from arrow_matrix_b200.arrow_matrix import ArrowMatrix
from arrow_matrix_b200.arrow_slim_mpi import ArrowSlimMPI
from tests.golden_util import GOLDEN_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "arrow_b200.h")).read()
    declared = set(re.findall(r"\b(arrow_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = _lib.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libarrow_b200.so does not export {name}"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    header_abi = int(re.search(r"#define ARROW_ABI_VERSION (\d+)", hdr).group(1))
    assert lib.arrow_b200_abi_version() == header_abi == _lib.ABI_VERSION


@pytest.mark.skipif(_cuda(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(_lib.ArrowError) as e:
        _lib.Context(0)
    assert "no CPU fallback" in str(e.value)
    from arrow_matrix_b200.engine import ArrowEngine
    dec = synth.synth_decomposition(2, 8, levels=1)
    with pytest.raises(_lib.ArrowError):
        ArrowEngine(dec, 8, 4)


def test_device_cpu_is_refused():
    from arrow_matrix_b200.comm import SelfComm
    with pytest.raises(NotImplementedError):
        ArrowDecompositionMPI.initialize(SelfComm(), np.array([2, 1]), None, None, 8, 4, device='cpu')
    with pytest.raises(NotImplementedError):
        ArrowSlimMPI(SelfComm()).spmm(device='cpu')


def test_surface_has_reference_methods():
    for name in ["result_tile", "feature_tile", "spmm", "set_features", "load_sparse_matrix_from_blocks",
                 "is_column_rank", "zero_rhs", "allgather_result", "set_features_slice_from_features"]:
        assert name in ArrowMatrix.__abstractmethods__
        assert callable(getattr(ArrowSlimMPI, name))
    for name in ["load_decomposition_new", "initialize", "step", "_propagate_features", "_aggregate",
                 "_all_to_all_tables", "number_of_blocks", "load_data_from_blocks"]:
        assert callable(getattr(ArrowDecompositionMPI, name))


def test_all_to_all_tables_product_version_against_reference_outputs():
    z = np.load(os.path.join(GOLDEN_DIR, "all_to_all_tables.npz"))
    for i in range(int(z["n"])):
        head = z[f"in_{i}"]
        rpr, cols, total, off = (int(x) for x in head[:4])
        c, d, sp, rp = ArrowDecompositionMPI._all_to_all_tables(head[4:], rpr, cols, total, off)
        assert np.array_equal(c, z[f"counts_{i}"]) and np.array_equal(d, z[f"displs_{i}"])
        assert np.array_equal(sp, z[f"send_{i}"]) and np.array_equal(rp, z[f"recv_{i}"])
    # the reference's own assertions (tests/test_arrowmpi.py:24-47)
    ranks, prev_ranks, rpr, cols = 2, 6, 4, 6
    perm = np.asarray(list(reversed(range(ranks * rpr))))
    for i in range(ranks):
        sl = perm[i * rpr:(i + 1) * rpr]
        counts, displs, p, out_p = ArrowDecompositionMPI._all_to_all_tables(sl, rpr, cols, prev_ranks + ranks, prev_ranks)
        assert counts[ranks + prev_ranks - i - 1] == rpr * cols and sum(counts) == rpr * cols
        assert displs[ranks + prev_ranks - i - 1] == 0


def test_load_decomposition_new_surface(tmp_path):
    from arrow_matrix_b200.comm import SelfComm
    dec = synth.synth_decomposition(4, 8, levels=2, perm_kind="random", seed=9)
    base = str(tmp_path / "g")
    graphio.save_decomposition_new(dec, base, 8, True)
    blocks, n_blocks, to_prev, to_next = ArrowDecompositionMPI.load_decomposition_new(SelfComm(), base, 8, True, slim=True)
    assert list(n_blocks) == [4, 2] and n_blocks.dtype == np.int32
    assert to_prev[0] is None and to_next[1] is None
    assert np.array_equal(to_prev[1], dec[1][1])            # level 0 is the identity
    assert blocks.width == 8 and len(blocks.decomposition) == 2
    # missing files: same "nothing found" signalling as the reference (None + empty n_blocks)
    b2, nb2, _, _ = ArrowDecompositionMPI.load_decomposition_new(SelfComm(), str(tmp_path / "nope"), 8, True)
    assert b2 is None and nb2.size == 0


def test_cli_flags_match_reference():
    from arrow_matrix_b200 import cli
    import argparse
    with pytest.raises(SystemExit):
        cli.main(["--help"])
    assert cli.str2bool("yes") and not cli.str2bool("0")


def test_wb_logging_writes_the_reference_artefacts(tmp_path, monkeypatch):
    """same four files and the same content layout as the reference's file logger (wb_logging.py:81-114, 191-201)"""
    import pickle
    from arrow_matrix_b200 import wb_logging
    from arrow_matrix_b200.comm import SelfComm
    monkeypatch.chdir(tmp_path)
    assert wb_logging.wandb_init(SelfComm(), "data/graphs/toy", 16, 3, "gpu", "Arrow_B200_v0.1_Slim", 100) is None
    wb_logging.log({"init_time": 0.5})
    for i in range(3):
        wb_logging.set_iteration_data({"iteration": i})
        wb_logging.log({"spmm_time": 0.1 * (i + 1)})
    base = wb_logging.finish()
    assert base.startswith("logs/Arrow_B200_v0.1_Slim.toy.")
    with open(base + ".pickle", "rb") as f:
        data = pickle.load(f)
    assert data[0] == {"init_time": 0.5, "rank": 0}
    assert data[3] == {"spmm_time": 0.1 * 3, "iteration": 2, "rank": 0}
    assert open(base + ".txt").read() == str(data)
    with open(base + ".config.pickle", "rb") as f:
        cfg = pickle.load(f)
    assert cfg == {"dataset": "toy", "width": 100, "n_features": 16, "iterations": 3, "device": "gpu", "ranks": 1,
                   "host": "NA", "algorithm": "Arrow_B200_v0.1_Slim"}
    assert open(base + ".config").read() == str(cfg)
    runs = list(wb_logging.load_local_runs(tmp_path / "logs"))
    assert len(runs) == 1 and runs[0][0] == cfg and runs[0][1] == data
    open(base + ".logged", "w").close()                      # the reference marks uploaded runs this way
    assert list(wb_logging.load_local_runs(tmp_path / "logs")) == []
    wb_logging.wandb_init(SelfComm(), None, 4, 1, "gpu", "X", 10)
    assert wb_logging.logs() == [] and wb_logging._CONFIG["dataset"] == "synthetic"


def test_plain_c_caller_binds_the_abi(tmp_path):
    """include/arrow_b200.h compiles as strict C99 and a C program drives the library through dlopen; without a GPU
    the library refuses to work instead of falling back to the CPU"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    _lib.load_library()                                            # builds the library if it is missing
    so = os.path.join(ROOT, "arrow_matrix_b200", "libarrow_b200.so")
    exe = str(tmp_path / "c_abi_caller")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-o", exe,
                    os.path.join(ROOT, "tests", "c_abi_caller.c"), "-ldl"], check=True)
    out = subprocess.run([exe, so], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == ("gpu 36.0" if _cuda() else "no-gpu")


def test_npz_layout_roundtrip_and_old_naming(tmp_path):
    """the reference's .npz level files (graphio.py:73-117, 194-249): current naming, the old naming fallback, and the
    loader route of ``--npy false`` (arrow_dec_mpi.py:641-648)"""
    from scipy import sparse
    from arrow_matrix_b200.comm import SelfComm
    dec = synth.synth_decomposition(5, 8, levels=2, seed=4)
    base = str(tmp_path / "g")
    graphio.save_decomposition(dec, base, 8, block_diagonal=True)
    assert os.path.exists(base + "_B_8_0_bd.npz") and os.path.exists(base + "_B_8_0_bd_nnzrows.npy")
    back = graphio.load_decomposition(base, 8, True)
    assert len(back) == 2
    for (B, p), (B2, p2) in zip(dec, back):
        assert abs(sparse.csr_matrix(B) - B2).nnz == 0 and np.array_equal(p, p2)
    blocks, n_blocks, to_prev, to_next = ArrowDecompositionMPI.load_decomposition_new(SelfComm(), base, 8, True, use_npy=False)
    ref = ArrowDecompositionMPI.load_decomposition_new(SelfComm(), _npy_twin(dec, tmp_path), 8, True)
    assert list(n_blocks) == list(ref[1])
    assert all(np.array_equal(a, b) for a, b in zip(to_prev[1:], ref[2][1:]))
    # old naming: {base}_B_{width}_{i}_bd.npz
    old = str(tmp_path / "old")
    for i, (B, p) in enumerate(dec):
        sparse.save_npz(f"{old}_B_8_{i}_bd.npz", sparse.csr_matrix(B))
        np.save(f"{old}_B_8_{i}_bd_permutation.npy", p)
    back = graphio.load_decomposition(old, 8, True)
    assert len(back) == 2 and abs(back[1][0] - sparse.csr_matrix(dec[1][0])).nnz == 0
    assert graphio.load_decomposition(str(tmp_path / "missing"), 8, True) == []


def _npy_twin(dec, tmp_path):
    base = str(tmp_path / "twin")
    graphio.save_decomposition_new(dec, base, 8, True)
    return base


def test_wide_operator_class_is_handed_out():
    """initialize(slim=False) returns the reference's wide operator type (arrow_dec_mpi.py:166-197), slim=True the slim one"""
    from arrow_matrix_b200.arrow_mpi import ArrowMPI
    from arrow_matrix_b200.comm import SelfComm
    wide = ArrowDecompositionMPI.initialize(SelfComm(), np.array([3, 2]), None, None, 8, 4, 'gpu', False, False)
    assert type(wide.B) is ArrowMPI and wide.B.is_block_diagonal is False and len(wide.levels) == 2
    assert all(type(lv) is ArrowMPI and lv._owner is wide for lv in wide.levels)
    slim = ArrowDecompositionMPI.initialize(SelfComm(), np.array([3, 2]), None, None, 8, 4, 'gpu', True, True)
    assert type(slim.B) is ArrowSlimMPI and slim.levels[1]._level == 1
    assert issubclass(ArrowMPI, ArrowMatrix) and ArrowMPI(SelfComm()).is_block_diagonal is False
    with pytest.raises(RuntimeError):
        wide.B.spmm()                                   # blocks not loaded yet


def test_utils_module_keeps_reference_helpers():
    from scipy import sparse
    from arrow_matrix_b200 import utils
    assert utils.str2bool("yes") is True and utils.str2bool("0") is False and utils.time_to_ms(0.0125) == 12
    rng = np.random.default_rng(0)
    g = sparse.random(9, 9, density=0.3, format="csr", random_state=1, dtype=np.float64)
    perm = rng.permutation(9)
    out = utils.relabel_nodes(g, {int(i): int(p) for i, p in enumerate(perm)})
    dense = np.zeros((9, 9))
    d = g.toarray()
    for a in range(9):
        for b in range(9):
            dense[perm[a], perm[b]] = d[a, b]
    assert np.array_equal(out.toarray(), dense) and out.has_canonical_format
    with pytest.raises(ValueError):
        utils.relabel_nodes(g, {i: 0 for i in range(9)})
    with pytest.raises(TypeError):
        utils.relabel_nodes(g.tocsc(), {i: i for i in range(9)})
    A = utils.generate_sparse_matrix(50, 70, 500, np.float32, rng)
    X = utils.generate_dense_matrix(70, 3, np.float32, rng)
    assert A.shape == (50, 70) and A.has_canonical_format and X.min() >= -1 and X.max() < 1


@pytest.mark.skipif(_cuda(), reason="checks the no-GPU behaviour of the documented stub")
def test_integration_md_stub_binds_the_library():
    """the reference-side ctypes stub printed in INTEGRATION.md section 2 is real code: it binds every symbol it names
    with the documented signatures and reports the library's error text"""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text.split("```python")[1].split("```")[0]
    so = os.path.join(ROOT, "arrow_matrix_b200", "libarrow_b200.so")
    _lib.load_library()
    block = block.replace('ctypes.CDLL("libarrow_b200.so")', f'ctypes.CDLL({so!r})')
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    assert {"B200", "_ck", "_L"} <= set(ns)
    with pytest.raises(RuntimeError) as e:
        ns["B200"](0)
    assert "no CPU fallback" in str(e.value)
