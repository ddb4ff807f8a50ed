"""``spmm_petsc`` command line -- same flags as the reference's ``scripts/spmm_petsc_main.py:9-45``."""
import argparse
import os

import numpy as np

from ..cli import str2bool
from .spmm_petsc import benchmark_spmm


def main(argv=None) -> None:
    parser = argparse.ArgumentParser(description='SpMM PETSc-style (1D halo exchange) benchmark on B200.')
    parser.add_argument('-s', '--seed', type=int, nargs="?", default=42, help='The seed for the random number generator.')
    parser.add_argument('-t', '--type', nargs="?", choices=['float32', 'float64'], default='float32',
                        help='The type of the data (the device path computes in float32).')
    parser.add_argument('-f', '--file', type=str, nargs="?", default=None,
                        help='A slice of the sparse matrix, {name}.part.{x}.slice.{y}.npz for a partition into x parts (any y).')
    parser.add_argument('-c', '--columns', type=int, nargs="?", default=32, help='The number of columns in the matrix X.')
    parser.add_argument('-i', '--device', type=str, default='gpu', help='Only gpu here.')
    parser.add_argument('-z', '--iterations', type=int, default=3, help='Number of iterations to benchmark.')
    parser.add_argument('--gpu-tiling', type=str2bool, nargs="?", default=False, help='Accepted for compatibility; not needed.')
    parser.add_argument('--dryrun', type=str2bool, nargs="?", default=False, help='Build the tables only (no benchmark).')
    parser.add_argument('-m', '--memory', type=float, default=0.9, help='Accepted for compatibility.')
    args = vars(parser.parse_args(argv))
    from .. import comm as comm_mod
    comm_mod.init_from_env()                    # torchrun --nproc-per-node N: one process per GPU
    file = None if args['file'] in (None, "None") else args['file']
    benchmark_spmm(file, args['columns'], args['iterations'], args['device'], os.environ.get('WANDB_API_KEY'),
                   np.dtype(args['type']), np.random.default_rng(args['seed']), args['gpu_tiling'], args['dryrun'],
                   args['memory'])


if __name__ == '__main__':
    main()
