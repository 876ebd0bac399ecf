"""ORACLE / TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's message-passing hot path.  Nothing under meta_gcn_b200/ imports
this package; only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
bench.py do.  See oracle/port.py for the parity status.
"""
import os
import sys

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def use_shim():
    """Make the CPU restatements of torch_scatter / torch_geometric / torch_sparse importable under
    their third-party names (they are absent from this image)."""
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    return SHIM
