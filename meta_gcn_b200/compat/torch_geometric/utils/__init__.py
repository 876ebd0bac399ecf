"""``torch_geometric.utils`` names on the hot path: ``scatter_`` (graph_attention.py:5 imports it at
module load; PyG's MessagePassing uses it) and ``degree`` (kernel/datasets.py:16)."""
from .... import functional as F_mgcn
from .... import ops
from ....graph import structure_of_index


def scatter_(name, src, index, dim_size=None):
    """PyG 1.3 utils.scatter_: 'add' | 'mean' | 'max' along dim 0; for 'max' the fill value (-1e9) of rows without
    entries is replaced by 0, which is what the first-maximum kernel writes there"""
    assert name in ["add", "mean", "max"]
    if name == "max":
        return F_mgcn.scatter_rows_max(src, index, dim_size)[0]
    return F_mgcn.scatter_rows(src, index, dim_size, name)


def degree(index, num_nodes=None, dtype=None):
    if num_nodes is None:
        num_nodes = int(index.max().item()) + 1 if index.numel() else 0
    deg = ops.degree_impl(structure_of_index(index, num_nodes).fwd.rowptr)
    return deg if dtype is None else deg.to(dtype)
