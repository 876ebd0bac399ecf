"""ORACLE ONLY — torch_sparse.utils.unique, the single torch_sparse symbol on the path
(data_procs/undirected.py:2)."""
