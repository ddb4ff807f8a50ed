"""``spmm_15d`` command line -- same flags as the reference's ``scripts/spmm_15d_main.py:20-75``."""
import argparse
import os

import numpy as np
from scipy import sparse

from .. import synth
from ..cli import str2bool
from .spmm_15d import benchmark_15d


def main(argv=None):
    parser = argparse.ArgumentParser(description='SpMM 1.5D (A-stationary, replication factor c) benchmark on B200.')
    parser.add_argument('-d', '--dataset', nargs="?", choices=['random', 'file'], default='random')
    parser.add_argument('-s', '--seed', type=int, nargs="?", default=42)
    parser.add_argument('-v', '--vertices', type=int, nargs="?", default=100000)
    parser.add_argument('-e', '--edges', type=int, nargs="?", default=1000000)
    parser.add_argument('-t', '--type', nargs="?", choices=['float32', 'float64'], default='float32')
    parser.add_argument('-f', '--file', type=str, nargs="?", default=None,
                        help='.npz adjacency matrix (old) or the prefix of _indptr/_indices/_data .npy files (new)')
    parser.add_argument('-c', '--columns', type=int, nargs="?", default=128)
    parser.add_argument('-r', '--replication', type=int, nargs="?", default=0,
                        help='0: the largest power of two whose square is at most the number of GPUs')
    parser.add_argument('--validate', type=str2bool, nargs="?", default=True)
    parser.add_argument('-i', '--device', type=str, default='gpu')
    parser.add_argument('-z', '--iterations', type=int, default=10)
    parser.add_argument('--gpu-tiling', type=str2bool, nargs="?", default=False, help='Accepted for compatibility.')
    parser.add_argument('-y', '--decomposition', type=str, default='old', help='old (npz / generated) or new (npy triplet)')
    args = vars(parser.parse_args(argv))
    if args['type'] != 'float32':
        raise SystemExit("the device path computes in float32")
    from .. import comm as comm_mod
    comm_mod.init_from_env()
    comm = comm_mod.world_comm()
    rng = np.random.default_rng(args['seed'])
    new = args['decomposition'] != 'old'
    A = None
    if args['dataset'] == 'file':
        if args['file'] is None:
            raise SystemExit("Please specify the file containing the adjacency matrix.")
        path = os.path.abspath(args['file'])
        if new:                                 # every rank maps the files and cuts its own block
            A = tuple(np.lib.format.open_memmap(f"{path}_{part}.npy", mode='r') for part in ("data", "indices", "indptr"))
        elif comm.Get_rank() == 0:
            A = sparse.load_npz(path).astype(np.float32)
    elif comm.Get_rank() == 0:
        A = synth.generate_sparse_matrix(args['vertices'], args['vertices'], args['edges'], np.float32, rng)
    return benchmark_15d(A, args['columns'], args['replication'], args['iterations'], args['device'], rng,
                         validate=args['validate'], new_decomposition=new,
                         dataset_name=args['file'] if args['dataset'] == 'file' else 'random', comm=comm,
                         wandb_api_key=os.environ.get('WANDB_API_KEY'))


if __name__ == '__main__':
    main()
