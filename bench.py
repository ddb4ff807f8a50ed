#!/usr/bin/env python
"""Benchmark of the iterated arrow-decomposed SpMM hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the B200 engine
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU arithmetic on host cores

A step is one ``ArrowDecompositionMPI.step()`` (forward exchange -> per-level arrow SpMM -> backward
scatter-add) over the synthetic decomposition G2 of SURVEY.md 8d: 10M rows, width 10 000, two levels,
~10 nnz/row, k = 128 fp32 features, uniformly random level-1 permutation (seed 503).  Prints ONE JSON line.

* ``value``      GFLOP/s = 2 * sum(nnz) * k / time, features and matrices resident in HBM, CUDA-event timed
* ``e2e``        same metric through the reference-facing classes with HOST buffers: every step uploads the
                 features from pinned memory and downloads the result tile
* ``roofline``   level-0 arrow SpMM kernel alone: algorithmic bytes / CUDA-event time vs the measured HBM peak
* ``cpu_baseline`` the reference's CPU path (oracle port of SciPy's kernel on host threads), bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=1000, help="block-rows of level 0 (x width = rows)")
    ap.add_argument("--width", type=int, default=10000)
    ap.add_argument("--k", type=int, default=128)
    ap.add_argument("--levels", type=int, default=2)
    ap.add_argument("--perm", type=str, default="random", choices=["random", "local", "identity"])
    ap.add_argument("--mode", type=str, default="auto", choices=["auto", "fused", "exchange"])
    ap.add_argument("--exchange", type=str, default="p2p", choices=["p2p", "p2p-direct", "nccl"],
                    help="multi-GPU level exchange: NVLink peer pulls (default) or NCCL all-to-all")
    ap.add_argument("--overlap", type=int, default=1, help="multi-GPU: 0 = serial phases, 1 = forward exchange overlaps the level-0 SpMM (default), 2 = both exchanges overlap a split level-0 SpMM (two levels, p2p exchange; not yet run on hardware)")
    ap.add_argument("--l2-hints", type=str, default="", help="plain,fused L2 hint masks of the tile kernel (e.g. 3,0)")
    ap.add_argument("--prefetch", type=int, default=-1, help="tile kernel L2 prefetch mask (bit0 plain, bit1 fused); -1 = library default")
    ap.add_argument("--fused-style", type=str, default="gather", choices=["gather", "scatter"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the full-size parity property (one step on all-ones features)")
    ap.add_argument("--cpu-sample-blocks", type=int, default=0)
    return ap.parse_args()


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        load = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_name(a):
    return (f"G2 synthetic arrow decomposition: {a.blocks * a.width} rows, width {a.width}, {a.levels} levels, "
            f"~10 nnz/row, k={a.k} fp32, level-1 permutation {a.perm} (seed 503)")


def build_decomposition(a, blocks=None):
    from arrow_matrix_b200 import synth
    return synth.synth_decomposition(blocks or a.blocks, a.width, levels=a.levels, perm_kind=a.perm, seed=503)


def traffic_from_profile(a):
    """dram bytes per launch of the level-0 SpMM from the committed ncu capture (profiles/), if it matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        key = f"blocks{a.blocks}_w{a.width}_k{a.k}"
        return t.get(key)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------------
def run_reference(a):
    """The reference's CPU implementation of the path on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_parallel
    from arrow_matrix_b200 import synth
    cores = os.cpu_count() or 1
    blocks = a.cpu_sample_blocks or (a.blocks if cores >= 64 else min(a.blocks, 250) if cores >= 16 else min(a.blocks, 100))
    dec = build_decomposition(a, blocks)
    ref = cpu_parallel.CpuArrowReference(dec, a.width, a.k, n_threads=cores)
    rng = np.random.default_rng(42)
    X = synth.generate_dense_matrix(ref.rows[0], a.k, np.float32, rng)
    for _ in range(a.warmup):
        ref.set_features(X)
        ref.step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ref.set_features(X)
        ref.step()
    dt = (time.perf_counter() - t0) / a.steps
    gflops = ref.flops_per_step() / dt / 1e9
    sample = (f"{blocks} of {a.blocks} block-rows of the same generator ({blocks * a.width} rows), full step "
              f"(gather, 2 products, scatter-add), {a.steps} timed steps")
    line = {"impl": "reference", "metric": "iterated SpMM GFLOP/s (k=%d)" % a.k, "value": gflops, "unit": "GFLOP/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "cpu_path": "oracle port of scipy csr_matvecs + row gather/scatter-add on host threads "
                       "(the reference's arithmetic; the literal reference needs mpi4py and >= 1500 MPI ranks)"},
            "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)



# ----------------------------------------------------------------------------------------------------------
# full-size parity property: one step on all-ones features
# ----------------------------------------------------------------------------------------------------------
def expected_ones_step(decomposition, width, block_diagonal=True):
    """float64 column of ``C_0`` after ONE step on all-ones features (every feature column is the same): the row sums
    of every level's arrow blocks pushed through the exchange maps.  Host arithmetic on the CSR arrays only -- a
    size-independent checksum of the whole iteration (forward exchange, every product, backward scatter-add).
    Returns ``(expected, state_free)``; ``state_free`` is False when some row lies behind the sentinel (its value then
    depends on earlier iterations, arrow_dec_mpi.py:544) and the property does not apply."""
    from scipy import sparse
    from arrow_matrix_b200 import decomp
    L = len(decomposition)
    n_blocks = [decomp.number_of_blocks(B, width) for B, _ in decomposition]
    _, to_prev, _, _ = decomp.prepare_permutations([p for _, p in decomposition], n_blocks, width)
    rows = [int(b) * width for b in n_blocks]
    x = [np.ones(rows[0])]
    state_free = True
    for j in range(1, L):
        tp = to_prev[j][: rows[j]]
        valid = tp < rows[j - 1]
        state_free = state_free and bool(valid.all())
        x.append(np.where(valid, x[j - 1][np.where(valid, tp, 0)], 0.0))
    c = []
    for j, (B, _) in enumerate(decomposition):
        ip, idx, dat, _ = decomp.arrow_rows(B, width, n_blocks[j], block_diagonal, 0, rows[j])
        vals = np.ones(idx.size) if dat is None else np.asarray(dat, dtype=np.float64)
        c.append(sparse.csr_matrix((vals, idx, ip), shape=(rows[j], rows[j])) @ x[j])
    for j in range(L - 1, 0, -1):
        tp = to_prev[j][: rows[j]]
        valid = tp < rows[j - 1]
        c[j - 1][tp[valid]] += c[j][valid]                  # the maps are injective
    return c[0], state_free


def verify_ones_step(eng, decomposition, width, row0, hostX, hostC, comm, tol=1e-5):
    """Run the property at the benchmark's own size, outside every timed region.  Never raises: a failure of the
    check itself is reported in the JSON line instead of losing the measurement, and every rank takes part in the same
    collectives whatever happens locally (the step is collective at N > 1)."""
    name = "one step on all-ones features == row sums of every level pushed through the exchange maps"
    expected, state_free, problem = None, None, None
    try:
        expected, state_free = expected_ones_step(decomposition, width)
    except Exception as e:     # noqa: BLE001
        problem = f"{type(e).__name__}: {e}"
    try:
        state = comm.allgather((problem, state_free))
    except Exception as e:     # noqa: BLE001
        return {"property": name, "error": f"{type(e).__name__}: {e}"}
    if any(p for p, _ in state):
        return {"property": name, "error": next(p for p, _ in state if p)}
    if not all(sf for _, sf in state):
        return {"property": name, "skipped": "rows behind the sentinel make the result state dependent"}
    rel = float("nan")
    try:
        hostX.array[:] = 1.0
        eng.set_features(hostX.array)
        eng.step()
        got = eng.result(0, hostC.array)
        n = got.shape[0]
        exp = expected[row0: row0 + n]
        scale = max(float(np.max(np.abs(expected))), 1e-30)
        err = 0.0
        for a0 in range(0, n, 1 << 20):                      # chunks: no 10 GB float64 temporary
            a1 = min(n, a0 + (1 << 20))
            err = max(err, float(np.max(np.abs(got[a0:a1].astype(np.float64) - exp[a0:a1, None]))))
        rel = err / scale
    except Exception as e:     # noqa: BLE001
        problem = f"{type(e).__name__}: {e}"
    try:
        outcome = comm.allgather((problem, rel))
    except Exception as e:     # noqa: BLE001
        return {"property": name, "error": f"{type(e).__name__}: {e}"}
    if any(p for p, _ in outcome):
        return {"property": name, "error": next(p for p, _ in outcome if p)}
    worst = float(max(r for _, r in outcome))
    return {"property": name, "rows": int(expected.size), "max_rel_err": worst, "tolerance": tol, "ok": bool(worst <= tol)}


# ----------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 engine has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from arrow_matrix_b200 import _lib, graphio, synth
    from arrow_matrix_b200.arrow_dec_mpi import ArrowDecompositionMPI
    from arrow_matrix_b200 import comm as comm_mod

    t_setup = time.time()
    dec = build_decomposition(a)
    comm = comm_mod.world_comm()
    if world > 1:
        from arrow_matrix_b200.sharded import ShardedArrowDecomposition
        arrow = ShardedArrowDecomposition(comm, dec, a.width, a.k, device=local_rank, exchange=a.exchange,
                                          overlap=a.overlap)
        eng = arrow.engine
    else:
        # the public path: files on disk -> load_decomposition_new -> initialize -> load blocks
        base = os.path.join(ROOT, "tmp", f"bench_{a.blocks}_{a.width}_{a.levels}_{a.perm}")
        graphio.save_decomposition_new(dec, base, a.width, block_diagonal=True)
        del dec
        blocks, n_blocks, to_prev, to_next = ArrowDecompositionMPI.load_decomposition_new(comm, base, a.width, True, slim=True)
        arrow = ArrowDecompositionMPI.initialize(comm, n_blocks, to_prev, to_next, a.width, a.k, 'gpu', True, True, mode=a.mode)
        arrow._fused_style = a.fused_style
        arrow.B.load_sparse_matrix_from_blocks(blocks)
        arrow.B.zero_rhs(a.width, a.k)
        eng = arrow._engine
        dec = blocks.decomposition                 # memory-mapped level files (for the full-size property check)
    ctx = eng.ctx
    if a.l2_hints:
        hp, hf = (int(x) for x in a.l2_hints.split(","))
        ctx.set_option(ctx.OPT_L2_HINTS_PLAIN, hp)
        ctx.set_option(ctx.OPT_L2_HINTS_FUSED, hf)
    if a.prefetch >= 0:
        ctx.set_option(ctx.OPT_PREFETCH, a.prefetch)
    rows_local = eng.local_rows if hasattr(eng, "local_rows") else eng.levels[0].rows
    rng = np.random.default_rng(42 + rank)
    hostX = _lib.PinnedArray((rows_local, a.k))
    hostC = _lib.PinnedArray((rows_local, a.k))
    hostX.array[:] = 2 * rng.random((rows_local, a.k), dtype=np.float32) - 1
    eng.set_features(hostX.array)
    ctx.sync()
    t_setup = time.time() - t_setup

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.sync()

    # ---- device-resident throughput -------------------------------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        eng.rewind_features()
        eng.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    ctx.timer_start(0)
    for _ in range(a.steps):
        eng.rewind_features()
        eng.step()
    ctx.timer_stop(0)
    barrier()
    ms_total = ctx.timer_ms(0)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / a.steps
    if dist is not None:
        t = torch.tensor([ms_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    flops = eng.flops_per_step()
    alg_bytes = eng.algorithmic_bytes_per_step()

    # ---- dominant kernel alone: level-0 arrow SpMM -----------------------------------------------------------
    roof = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        kms = eng.time_level_spmm(0, a.steps) if hasattr(eng, "time_level_spmm") else None
        if kms:
            kb = eng.level_bytes(0)
            roof = {"bound": "hbm", "achieved": kb / kms / 1e6, "peak": peak, "unit": "GB/s", "frac": kb / kms / 1e6 / peak,
                    "traffic": traffic_from_profile(a), "kernel": "k_spmm_tiles level 0 (one launch%s)" % ("" if world == 1 else ", rank 0's shard"), "kernel_ms": kms,
                    "algorithmic_bytes_per_launch": kb, "peak_source": peak_src}

    # ---- end to end through the public classes with host buffers --------------------------------------------
    e2e = None
    if not a.no_e2e:
        n_e2e = max(4, min(a.steps, 10))
        nbytes = rows_local * a.k * 4
        streaming = hasattr(arrow, "step_stream")
        if streaming:
            # two (features, result) pairs of pinned host buffers in rotation; uploads / compute / downloads of
            # consecutive iterations overlap on copy lanes (PCIe is full duplex) -- every step still moves its
            # own 5.12 GB up and 5.12 GB down
            hx = [hostX, _lib.PinnedArray((rows_local, a.k))]
            hc = [hostC, _lib.PinnedArray((rows_local, a.k))]
            hx[1].array[:] = hostX.array
            for i in range(2):
                arrow.step_stream(hx[i % 2].array, hc[i % 2].array)
            arrow.synchronize()
            barrier()
            t0 = time.perf_counter()
            for i in range(n_e2e):
                arrow.step_stream(hx[i % 2].array, hc[i % 2].array)
            arrow.synchronize()
            barrier()
            dt = (time.perf_counter() - t0) / n_e2e
            api = "ArrowDecompositionMPI.step_stream(X_host, out_host) x N + synchronize() (pinned host buffers)"
        else:
            for _ in range(2):
                arrow.B.set_features(hostX.array)
                arrow.step()
                arrow.B.result_tile(out=hostC.array)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                arrow.B.set_features(hostX.array)       # pinned host -> device, inside the timed region
                arrow.step()
                arrow.B.result_tile(out=hostC.array)    # device -> pinned host (synchronises)
            barrier()
            dt = (time.perf_counter() - t0) / n_e2e
            api = "B.set_features / step / B.result_tile (blocking)"
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": flops / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes),
               "ms_per_step": dt * 1e3, "steps": n_e2e, "api": api}
        # blocking variant for the record (the reference's call sequence): one sample
        if streaming:
            t0 = time.perf_counter()
            arrow.B.set_features(hostX.array)
            arrow.step()
            arrow.B.result_tile(out=hostC.array)
            e2e["blocking_ms_per_step"] = (time.perf_counter() - t0) * 1e3

    # ---- CPU baseline (rank 0, bounded sample) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and not a.no_cpu and world == 1:
        from oracle import cpu_parallel
        cores = os.cpu_count() or 1
        sb = a.cpu_sample_blocks or min(a.blocks, 100 if cores < 32 else 250)
        sdec = build_decomposition(a, sb)
        ref = cpu_parallel.CpuArrowReference(sdec, a.width, a.k, n_threads=cores)
        ref.set_features(synth.generate_dense_matrix(ref.rows[0], a.k, np.float32, np.random.default_rng(1)))
        dt = ref.time_steps(3, warmup=1)
        cpu = {"value": ref.flops_per_step() / dt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
               "sample": f"{sb} of {a.blocks} block-rows of the same generator, 3 full steps on {cores} host threads "
                         f"({dt * 1e3:.0f} ms/step)"}
        ref.close()

    # ---- full-size parity property (untimed; all ranks take part in the step) ----------------------------------
    verified = None
    if not a.no_verify:
        row0 = eng.plan.levels[0].r0 if hasattr(eng, "plan") else 0
        verified = verify_ones_step(eng, dec, a.width, row0, hostX, hostC, comm)

    if rank == 0:
        line = {"metric": "iterated SpMM GFLOP/s (k=%d)" % a.k, "value": flops / ms_step / 1e6, "unit": "GFLOP/s",
                "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(a), "mode": eng.mode + ("/" + eng.fused_style if getattr(eng, "fused_style", None) and eng.mode == "fused" else ""), "overlap": (2 if getattr(eng, "split", False) else int(bool(getattr(eng, "overlap", False)))), "l2": "inputs larger than L2 (features 5.12 GB per pass at the default size); no flush",
                           "total_nnz": int(eng.total_nnz), "setup_s": round(t_setup, 1)},
                "hbm_gbs_effective": alg_bytes / ms_step / 1e6, "algorithmic_bytes_per_step": alg_bytes,
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "verified": verified}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
