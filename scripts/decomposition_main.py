#!/usr/bin/env python
"""arrow_decompose entry point (reference: scripts/decomposition_main.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_matrix_b200.decompose_cli import main  # noqa: E402

if __name__ == '__main__':
    main()
