"""Comparison baselines of the reference (``arrow/baseline/``) on the same C-ABI SpMM kernel."""
