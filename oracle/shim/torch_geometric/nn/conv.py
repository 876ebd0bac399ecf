"""ORACLE ONLY — CPU restatement of the PyG 1.3.x operators behind kernel/gcn.py:4,
kernel/gin.py:4, kernel/graph_sage.py:4 (SURVEY.md Appendix A is the spec; torch_geometric itself is
an un-vendored, un-pinned dependency of the reference, so this part of the oracle is anchored on the
reference's call sites and on the in-repo legacy copy src/gcn_meta/models/gcn.py:57-107).
"""
import inspect

import torch
from torch.nn import Parameter

from ..utils import add_remaining_self_loops, remove_self_loops, scatter_
from .inits import glorot, reset, uniform, zeros


class MessagePassing(torch.nn.Module):
    """propagate(): x_j = x[edge_index[0]] -> message -> scatter_(aggr, ., edge_index[1]) -> update
    (flow source_to_target).  Accepts both the 1.3 call `propagate(edge_index, size=None, **kw)` and
    the 1.0 call `propagate(aggr, edge_index, **kw)` used by src/gcn_meta/models/gcn.py:86."""

    def __init__(self, aggr="add", flow="source_to_target"):
        super().__init__()
        assert aggr in ["add", "mean", "max"] and flow == "source_to_target"
        self.aggr = aggr
        self._msg_args = [a for a in inspect.signature(self.message).parameters]
        self._upd_args = [a for a in inspect.signature(self.update).parameters][1:]

    def propagate(self, *args, size=None, **kwargs):
        if isinstance(args[0], str):
            aggr, edge_index = args[0], args[1]
        else:
            aggr, edge_index = self.aggr, args[0]
        num_nodes = None
        msg_in = []
        for name in self._msg_args:
            if name.endswith("_j") or name.endswith("_i"):
                t = kwargs[name[:-2]]
                num_nodes = t.size(0)
                sel = edge_index[0] if name.endswith("_j") else edge_index[1]
                msg_in.append(torch.index_select(t, 0, sel))
            else:
                msg_in.append(kwargs.get(name))
        if num_nodes is None:
            num_nodes = size if isinstance(size, int) else int(edge_index.max()) + 1
        out = self.message(*msg_in)
        out = scatter_(aggr, out, edge_index[1], dim_size=num_nodes)
        return self.update(out, *[kwargs.get(a) for a in self._upd_args])

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out


class GCNConv(MessagePassing):
    def __init__(self, in_channels, out_channels, improved=False, cached=False, bias=True):
        super().__init__("add")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached, self.cached_result = improved, cached, None
        self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.weight)
        zeros(self.bias)
        self.cached_result = None

    @staticmethod
    def norm(edge_index, num_nodes, edge_weight, improved=False, dtype=None):
        from torch_scatter import scatter_add
        if edge_weight is None:
            edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
        fill_value = 1 if not improved else 2
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill_value, num_nodes)
        row, col = edge_index
        deg = scatter_add(edge_weight, row, dim=0, dim_size=num_nodes)
        deg_inv_sqrt = deg.pow(-0.5)
        deg_inv_sqrt[deg_inv_sqrt == float("inf")] = 0
        return edge_index, deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]

    def forward(self, x, edge_index, edge_weight=None):
        x = torch.matmul(x, self.weight)
        if not self.cached or self.cached_result is None:
            self.cached_result = self.norm(edge_index, x.size(0), edge_weight, self.improved, x.dtype)
        edge_index, norm = self.cached_result
        return self.propagate(edge_index, x=x, norm=norm)

    def message(self, x_j, norm):
        return norm.view(-1, 1) * x_j

    def update(self, aggr_out):
        if self.bias is not None:
            aggr_out = aggr_out + self.bias
        return aggr_out


class SAGEConv(MessagePassing):
    """PyG 1.3: mean over (neighbours U self) then one weight + bias."""

    def __init__(self, in_channels, out_channels, normalize=False, bias=True):
        super().__init__("mean")
        self.in_channels, self.out_channels, self.normalize = in_channels, out_channels, normalize
        self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        uniform(self.in_channels, self.weight)
        uniform(self.in_channels, self.bias)

    def forward(self, x, edge_index, size=None):
        edge_index, _ = add_remaining_self_loops(edge_index, None, 1, x.size(0))
        return self.propagate(edge_index, size=size, x=x)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        aggr_out = torch.matmul(aggr_out, self.weight)
        if self.bias is not None:
            aggr_out = aggr_out + self.bias
        if self.normalize:
            aggr_out = torch.nn.functional.normalize(aggr_out, p=2, dim=-1)
        return aggr_out


class GINConv(MessagePassing):
    def __init__(self, nn, eps=0, train_eps=False):
        super().__init__("add")
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))
        self.reset_parameters()

    def reset_parameters(self):
        reset(self.nn)
        self.eps.data.fill_(self.initial_eps)

    def forward(self, x, edge_index):
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        edge_index, _ = remove_self_loops(edge_index)
        return self.nn((1 + self.eps) * x + self.propagate(edge_index, x=x))

    def message(self, x_j):
        return x_j
