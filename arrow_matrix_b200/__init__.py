"""arrow_matrix_b200 -- B200-native engine for the iterated arrow-decomposed SpMM hot path.

Public surface (mirrors spcl/arrow-matrix's class names for that path):

    from arrow_matrix_b200.arrow_dec_mpi import ArrowDecompositionMPI
    from arrow_matrix_b200.arrow_bench import bench_spmm

The arithmetic lives in ``libarrow_b200.so`` (``csrc/arrow_b200.cu``, C ABI in ``include/arrow_b200.h``);
there is no CPU fallback.
"""
__version__ = "0.1.0"
