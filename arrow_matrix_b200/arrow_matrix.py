"""Operator interface of one distributed arrow matrix -- same abstract surface as the reference's
``ArrowMatrix`` (``arrow/arrow_matrix.py:12-111``): the lower drop-in boundary of the hot path."""
from abc import ABC, abstractmethod
from typing import List

import numpy as np


class ArrowMatrix(ABC):
    # The number of tiles per side
    tiles_per_side: int

    @abstractmethod
    def result_tile(self):
        """Returns the result tile (this process's rows of C, host array)."""

    @abstractmethod
    def feature_tile(self):
        """Returns the feature tile (this process's rows of X, host array)."""

    @abstractmethod
    def spmm(self, device: str = 'gpu'):
        """Compute the SpMM of this level. ``device`` must be 'gpu' (there is no CPU path here)."""

    @abstractmethod
    def set_features(self, X: np.ndarray):
        """Sets this process's slice of the features (uploaded to the device)."""

    @abstractmethod
    def load_sparse_matrix_from_blocks(self, blocks):
        """Uploads the sparse blocks (once; they stay resident)."""

    @abstractmethod
    def is_column_rank(self) -> bool:
        """True if this process holds feature/result tiles."""

    @abstractmethod
    def zero_rhs(self, number_of_rows_per_rank: int, number_of_columns: int, dtype=np.float32):
        """Clears X and C; call before the first iteration."""

    @abstractmethod
    def allgather_result(self, C: np.ndarray):
        """Fills the caller-allocated ``(tiles_per_side*width) x k`` array with the whole level's result."""

    @abstractmethod
    def set_features_slice_from_features(self, X: np.ndarray):
        """Takes this process's rows out of the full feature matrix of the level."""
