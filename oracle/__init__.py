"""Test infrastructure: CPU restatement of the reference's hot path (see oracle/oracle.py)."""
