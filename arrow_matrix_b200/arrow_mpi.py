"""``ArrowMPI`` for B200: the reference's *wide* operator (``arrow/arrow_mpi.py:26-600``).

In the reference the wide layout spends ``2t-1`` MPI ranks on one arrow matrix (row tiles and column tiles) and is the
only layout that supports the banded mode (``A_i,i+-1`` blocks, ``:211-219, :262-269``).  On a GPU the rank layout
disappears: one launch per level multiplies every block of the arrow pattern, banded or not, so this class is the same
device operator as ``ArrowSlimMPI`` with the reference's constructor (``comm, is_block_diagonal``) and its
``is_block_diagonal`` attribute.  ``ArrowDecompositionMPI.initialize(..., slim=False)`` hands it out.
"""
from __future__ import annotations

from .arrow_slim_mpi import ArrowSlimMPI


class ArrowMPI(ArrowSlimMPI):
    def __init__(self, comm, is_block_diagonal: bool = False, owner=None, level: int = 0):
        super().__init__(comm, owner, level)
        self.is_block_diagonal = bool(is_block_diagonal)
