"""ORACLE ONLY — torch_geometric.utils (PyG 1.3) pieces used by the reference:
scatter_ (graph_attention.py:5), add_remaining_self_loops / remove_self_loops (inside GCNConv,
SAGEConv, GINConv), degree (kernel/datasets.py:16,58), add_self_loops (src/gcn_meta/models/gcn.py:5),
num_nodes.maybe_num_nodes (data_procs/undirected.py:3)."""
import torch
import torch_scatter

from .num_nodes import maybe_num_nodes


def scatter_(name, src, index, dim_size=None):
    assert name in ["add", "mean", "max"]
    op = getattr(torch_scatter, "scatter_{}".format(name))
    fill_value = -1e38 if name == "max" else 0
    out = op(src, index, 0, None, dim_size, fill_value)
    if isinstance(out, tuple):
        out = out[0]
    if name == "max":
        out[out == fill_value] = 0
    return out


def degree(index, num_nodes=None, dtype=None):
    num_nodes = maybe_num_nodes(index, num_nodes)
    out = torch.zeros((num_nodes,), dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, out.new_ones((index.size(0))))


def remove_self_loops(edge_index, edge_attr=None):
    row, col = edge_index
    mask = row != col
    edge_attr = edge_attr if edge_attr is None else edge_attr[mask]
    return edge_index[:, mask], edge_attr


def add_self_loops(edge_index, num_nodes=None):
    """PyG 1.0 form (returns only the index) used by the legacy src/gcn_meta/models/gcn.py:65."""
    num_nodes = maybe_num_nodes(edge_index, num_nodes)
    loop = torch.arange(0, num_nodes, dtype=torch.long, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop], dim=1)


def add_remaining_self_loops(edge_index, edge_weight=None, fill_value=1, num_nodes=None):
    num_nodes = maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index
    mask = row != col
    inv_mask = ~mask
    loop_weight = None
    if edge_weight is not None:
        assert edge_weight.numel() == edge_index.size(1)
        loop_weight = torch.full((num_nodes,), fill_value, dtype=edge_weight.dtype,
                                 device=edge_weight.device)
        remaining = edge_weight[inv_mask]
        if remaining.numel() > 0:
            loop_weight[row[inv_mask]] = remaining
        edge_weight = torch.cat([edge_weight[mask], loop_weight], dim=0)
    loop_index = torch.arange(0, num_nodes, dtype=row.dtype, device=row.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    edge_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
    return edge_index, edge_weight
