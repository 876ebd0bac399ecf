"""meta_gcn_b200 — B200 (sm_100a) message-passing engine behind meta-gcn's model API.

Layout: csrc/ (CUDA kernels + C-ABI, built to lib/libmgcn.so), ops.py (torch.library ops),
functional.py (autograd), graph.py (edge_index -> row structures), gcn_meta/ and kernel/ (mirrors of
the reference's model classes on this path), compat/ (torch_scatter / torch_geometric namespaces),
data.py (batching + synthetic workloads), dist.py (graph-sharded data parallelism).
"""
__version__ = "0.1.0"
