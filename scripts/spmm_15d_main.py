#!/usr/bin/env python
"""spmm_15d entry point (reference: scripts/spmm_15d_main.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_matrix_b200.baseline.spmm_15d_cli import main  # noqa: E402

if __name__ == '__main__':
    main()
