"""Where can the SpMM's row gathers come from, and how fast?  (libarrow_probes.so, csrc/probes.cu)

Prints bytes/s and bytes per clock per SM of the same gather loop fed from L2, from the CTA's own shared memory and
from the distributed shared memory of an 8-CTA cluster -- the evidence for / against a cluster-resident X panel.
"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arrow_matrix_b200 import build  # noqa: E402


def sm_clock_mhz():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True).stdout
        return float(out.strip().splitlines()[0])
    except Exception:
        return 1965.0


def main():
    lib = ctypes.CDLL(build.build_probes())
    lib.arrow_probe_gather.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double),
                                                           ctypes.POINTER(ctypes.c_int)]
    mhz = sm_clock_mhz()
    cases = [
        ("L2 -> SM, 512 B rows (k=128), 10000-row panel (5.12 MB)", 0, 512, 10000, 2048, 4),
        ("L2 -> SM, 128 B runs (k-slice 32), 10000-row panel", 0, 128, 10000, 4096, 4),
        ("L2 -> SM, 64 B rows (k=16), 10000-row panel", 0, 64, 10000, 4096, 4),
        ("own shared memory, 128 B runs, 1250-row slice (160 KB)", 1, 128, 1250, 8192, 1),
        ("cluster DSMEM (8 CTAs), 128 B runs, 10000-row panel over the cluster", 2, 128, 10000, 8192, 1),
    ]
    for name, mode, row_bytes, rows, iters, per_sm in cases:
        ms, nbytes, ctas = ctypes.c_float(), ctypes.c_double(), ctypes.c_int()
        best = None
        for _ in range(3):
            rc = lib.arrow_probe_gather(mode, row_bytes, rows, iters, per_sm, ctypes.byref(ms), ctypes.byref(nbytes), ctypes.byref(ctas))
            if rc != 0:
                break
            best = ms.value if best is None else min(best, ms.value)
        if best is None:
            print(json.dumps({"probe": name, "error": rc}), flush=True)
            continue
        tbs = nbytes.value / best / 1e9
        sms = 148
        print(json.dumps({"probe": name, "ms": round(best, 4), "TB_per_s": round(tbs, 2),
                          "B_per_clk_per_SM": round(nbytes.value / (best * 1e-3) / (mhz * 1e6) / sms, 1),
                          "ctas": ctas.value, "sm_mhz_assumed": mhz}), flush=True)


if __name__ == "__main__":
    main()
