"""Helpers to read the golden fixtures produced by tests/golden/make_golden.py (real reference runs)."""
import glob
import os

import numpy as np
from scipy import sparse

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# arrow-decomposition cases only: slim_* / wide_* (petsc_*.npz belong to the 1D baseline, all_to_all_tables.npz to the tables)
CASES = sorted(os.path.basename(p)[:-4] for pattern in ("slim_*.npz", "wide_*.npz")
               for p in glob.glob(os.path.join(GOLDEN_DIR, pattern)))


# every reference-generated fixture runs through the CUDA engine (tests/test_gpu_engine.py)
GPU_CASES = list(CASES)


class GoldenCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.width, self.k = int(z["width"]), int(z["k"])
        self.slim, self.block_diagonal = bool(z["slim"]), bool(z["block_diagonal"])
        self.L, self.iterations = int(z["levels"]), int(z["iterations"])
        self.n_blocks = [int(x) for x in z["n_blocks"]]
        self.one_based = bool(z["one_based"])
        self.write_data = bool(z["write_data"])
        self.decomposition = []
        for j in range(self.L):
            ip, ix, dt = z[f"indptr_{j}"], z[f"indices_{j}"], z[f"data_{j}"]
            n = ip.size - 1
            B = sparse.csr_matrix((dt, ix, ip), shape=(n, n))
            perm = z[f"perm_{j}"].astype(np.int64)
            self.decomposition.append((B, perm + 1 if self.one_based else perm))   # as stored in the files
        self.X = [z[f"X_{it}"] if bool(z[f"has_X_{it}"]) else None for it in range(self.iterations)]
        self.C = [[z[f"C_{it}_{j}"] for j in range(self.L)] for it in range(self.iterations)]
        self.final = [z[f"final_{j}"] for j in range(self.L)]
        self.to_prev = [z[f"to_prev_{j}"] for j in range(self.L)]
        self.to_next = [z[f"to_next_{j}"] for j in range(self.L)]
