"""Build recipe for libarrow_b200.so (in-tree, sm_100a only).

``python -m arrow_matrix_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a
GPU; the resulting .so sits next to this file so it travels with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "arrow_b200.cu")
HDR = os.path.join(os.path.dirname(HERE), "include", "arrow_b200.h")
OUT = os.path.join(HERE, "libarrow_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in (SRC, HDR, os.path.abspath(__file__)))


PROBES_SRC = os.path.join(HERE, "csrc", "probes.cu")
PROBES_OUT = os.path.join(HERE, "libarrow_probes.so")


def build_probes(force: bool = False) -> str:
    """libarrow_probes.so: measurement-only microbenchmarks (scripts/probe_gather.py); never loaded by the product."""
    if not force and os.path.exists(PROBES_OUT) and os.path.getmtime(PROBES_OUT) >= os.path.getmtime(PROBES_SRC):
        return PROBES_OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = f"{PROBES_OUT}.tmp.{os.getpid()}"
    res = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", tmp, PROBES_SRC], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libarrow_probes.so")
    os.replace(tmp, PROBES_OUT)
    return PROBES_OUT


def can_build() -> bool:
    import shutil
    return shutil.which(os.environ.get("NVCC", "nvcc")) is not None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile under an exclusive file lock into a temporary file and rename it into place: several processes
    (the ranks of one ``torchrun`` on a fresh checkout) may call this at once -- one compiles, the others wait for
    the lock, find the library up to date and return; nobody ever dlopens a half-written file."""
    import fcntl
    if not force and not needs_build():
        return OUT
    with open(OUT + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():           # another process built it while we waited
                return OUT
            nvcc = os.environ.get("NVCC", "nvcc")
            tmp = f"{OUT}.tmp.{os.getpid()}"
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, SRC]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libarrow_b200.so")
            os.replace(tmp, OUT)
            if verbose:
                sys.stderr.write(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_probes(force="--force" in sys.argv))
