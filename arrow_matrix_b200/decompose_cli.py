"""``arrow_decompose`` command line -- the reference's ``scripts/decomposition_main.py:109-208`` without igraph.

Same layout convention: the input lives at ``{dataset_dir}/{name}/{name}.mtx`` (Matrix Market, ``--format mtx``),
``.npz`` (scipy sparse, ``--format npz``) or ``.mat`` (``--format matlab``, MATLAB <= v7.2 through scipy.io; v7.3
files need the reference's mat73 path), the decomposition is written next to it as
``{name}_B_{width}_{i}_bd_{indptr,indices,data,permutation}.npy`` (block diagonal, at most 10 levels, like the
reference).  ``--visualize`` (matplotlib) is not provided."""
import argparse
from pathlib import Path

import numpy as np
from scipy import io as sio
from scipy import sparse

from . import graphio
from .decomposition import arrow_decomposition


def load_matrix(path: Path, fmt: str, directed: bool = False) -> sparse.csr_matrix:
    if fmt == "mtx":
        A = sparse.csr_matrix(sio.mmread(str(path)))
    elif fmt == "npz":
        A = sparse.csr_matrix(sparse.load_npz(str(path)))
    elif fmt == "matlab":
        mat = sio.loadmat(str(path))
        prob = mat["Problem"]
        A = sparse.csr_matrix(prob["A"][0, 0])
    else:
        raise ValueError(f"Unknown format {fmt}")
    if A.shape[0] != A.shape[1]:
        raise ValueError("the graph matrix must be square")
    A = A.astype(np.float32)
    if not directed:                       # undirected graph: symmetrise the pattern like igraph's undirected build
        A = A.maximum(A.T)
    A.setdiag(0)
    A.eliminate_zeros()
    A.data[:] = 1.0                        # adjacency, like graph.get_adjacency_sparse() (graphio.py:174)
    return A


def main(argv=None) -> None:
    parser = argparse.ArgumentParser(description="Arrow-decompose a sparse graph matrix into the npy layout")
    parser.add_argument('--width', type=int, default=5000000)
    parser.add_argument('--dataset_dir', type=str, default='~/Desktop/')
    parser.add_argument('--dataset_name', nargs='+', type=str, default=['kmer_V2a'])
    parser.add_argument('--format', type=str, default='mtx', help="mtx, npz or matlab")
    parser.add_argument('--directed', type=bool, default=False)
    parser.add_argument('--levels', type=int, default=10, help="maximum number of levels (reference: 10)")
    parser.add_argument('--seed', type=int, default=0)
    args = parser.parse_args(argv)
    print(args)
    if args.width <= 0:
        raise ValueError("Width must be positive")
    root = Path(args.dataset_dir).expanduser()
    ext = {"mtx": ".mtx", "npz": ".npz", "matlab": ".mat"}[args.format]
    for name in args.dataset_name:
        d = root / name
        f = d / (name + ext)
        if not f.exists():
            raise ValueError(f"File {f.name} does not exist in {d}")
        print(f"Loading {name}'s graph...")
        A = load_matrix(f, args.format, args.directed)
        print(f"Converting {name} ({A.shape[0]} vertices, {A.nnz} entries) to arrow decomposition with width {args.width}...")
        B = arrow_decomposition(A, arrow_width=min(args.width, A.shape[0]), max_number_of_levels=args.levels,
                                block_diagonal=True, seed=args.seed)
        print(f"Successfully decomposed into {len(B)} matrices.")
        graphio.save_decomposition_new(B, str(d / name), args.width, block_diagonal=True)
        print(f"Saved {d / name}_B_{args.width}_*_bd_*.npy")


if __name__ == "__main__":
    main()
