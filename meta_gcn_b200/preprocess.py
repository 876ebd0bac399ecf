"""GPU edge preprocessing with the reference's ordering — the step immediately before the hot path
(SURVEY.md §8 f1).  Mirrors of data_procs/undirected.py (``sort_unique_edges``, ``to_undirected_ey``),
data_procs/loop.py (``add_self_loops_ey``) and the out-degree feature of data_procs/data_add_degree.py:45-65,
all as ONE libmgcn call (two stable radix sorts, adjacent-unique, compaction, loop append, degree count)."""
import torch

from . import ops


def sort_unique_edges(edge_index, num_nodes):
    """data_procs/undirected.py:6-16: (edge_index sorted by row*N+col without duplicates, perm) where perm
    indexes one (the first) occurrence of each kept edge."""
    out, _deg, perm = ops.preprocess_edges_impl(edge_index, num_nodes, undirected=False, add_loops=False)
    return out, perm.long()


def to_undirected_ey(edge_index, edge_y=None, num_nodes=None):
    """data_procs/undirected.py:19-35."""
    n = int(edge_index.max().item()) + 1 if num_nodes is None else int(num_nodes)
    out, _deg, perm = ops.preprocess_edges_impl(edge_index, n, undirected=True, add_loops=False)
    if edge_y is not None:
        edge_y = torch.cat([edge_y, edge_y], dim=0)[perm.long()]
    return out, edge_y


def add_self_loops_ey(edge_index, edge_y=None, node_y=None, num_nodes=None):
    """data_procs/loop.py:5-24 (concatenation only: plumbing)."""
    n = int(edge_index.max().item()) + 1 if num_nodes is None else int(num_nodes)
    loop = torch.arange(n, dtype=torch.long, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    out = torch.cat([edge_index, loop], dim=1)
    if edge_y is not None:
        assert node_y is not None and len(node_y) == n
        edge_y = torch.cat([edge_y, node_y], dim=0)
    return out, edge_y


def botnet_edges(edge_index, num_nodes):
    """The whole data_procs/data_pre.sh:11-16 edge pipeline in one call: undirected + sort-unique + self loops
    appended + out-degree (the ``x[:,1]`` feature).  Returns (edge_index, deg)."""
    out, deg, _perm = ops.preprocess_edges_impl(edge_index, num_nodes, undirected=True, add_loops=True)
    return out, deg
