"""``torch_geometric.nn`` names imported by kernel/gcn.py:4, gin.py:4, graph_sage.py:4, with PyG 1.3
semantics, each forward a handful of libmgcn launches:

GCNConv   x·W -> add_remaining_self_loops -> deg over row -> d^-½[row]·d^-½[col] -> add -> +bias
SAGEConv  add_remaining_self_loops -> mean over (neighbours ∪ self) -> ·W + b
GINConv   remove_self_loops -> nn((1+eps)·x + Σ_j x_j)
"""
import torch
from torch.nn import Parameter

from .... import functional as F_mgcn
from .... import ops
from ....graph import LOOPS_ADD_REMAINING, LOOPS_REMOVE, structure_of
from . import inits
from .inits import glorot, reset, uniform, zeros


class GCNConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, improved=False, cached=False, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.weight)
        zeros(self.bias)

    def forward(self, x, edge_index, edge_weight=None, _act=None):
        n = x.size(0)
        graph = structure_of(edge_index, n, LOOPS_ADD_REMAINING)
        fill = 2.0 if self.improved else 1.0
        if edge_weight is None and not self.improved:
            deg = graph.out_degree()
        else:
            # PyG's add_remaining_self_loops keeps the weight of a self loop that is already in edge_index and gives
            # `fill` only to the nodes without one; this structure drops existing loops and appends one of weight `fill`
            # per node, which is the same thing only when there are none (kernel/gcn.py's unweighted default never
            # comes here).  Refuse the other case instead of returning different numbers.
            if graph.has_self_loops():
                raise NotImplementedError("GCNConv with improved=True or edge_weight on an edge_index that already "
                                          "contains self loops (their weights would have to be carried over)")
            ew = edge_weight if edge_weight is not None else torch.ones(
                edge_index.size(1), dtype=x.dtype, device=x.device)
            edge_weight = ew
            deg = graph.weighted_out_degree(ew, fill)
        dis = ops.gcn_norm_impl(deg, 0)
        xw = F_mgcn.linear(x, self.weight)
        return F_mgcn.aggregate(xw, graph, dis, dis, edge_weight, "add", self.bias, None, _act,
                                loop_value=fill)

    def __repr__(self):
        return "{}({}, {})".format(self.__class__.__name__, self.in_channels, self.out_channels)


class SAGEConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, normalize=False, bias=True):
        super().__init__()
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

    def forward(self, x, edge_index, size=None, _act=None):
        graph = structure_of(edge_index, x.size(0), LOOPS_ADD_REMAINING)
        mean = F_mgcn.aggregate(x, graph, reduce="mean")
        if self.normalize:
            out = F_mgcn.linear(mean, self.weight, self.bias)
            out = torch.nn.functional.normalize(out, p=2, dim=-1)
            return torch.relu(out) if _act == "relu" else out
        return F_mgcn.linear(mean, self.weight, self.bias, act=_act)

    def __repr__(self):
        return "{}({}, {})".format(self.__class__.__name__, self.in_channels, self.out_channels)


class GINConv(torch.nn.Module):
    def __init__(self, nn, eps=0, train_eps=False):
        super().__init__()
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
        graph = structure_of(edge_index, x.size(0), LOOPS_REMOVE)
        return self._mlp((1 + self.eps) * x + F_mgcn.aggregate(x, graph))

    def _mlp(self, h):
        """self.nn on libmgcn when it is the reference's Sequential of Linear / ReLU / BatchNorm1d (kernel/gin.py:10-16):
        each Linear (+ the ReLU behind it) is one mgcn_linear launch, BatchNorm1d the fixed-order kernels of
        csrc/batchnorm.cu.  Any other module is called as is."""
        mods = list(self.nn) if isinstance(self.nn, torch.nn.Sequential) else None
        if mods is None or not all(isinstance(m, (torch.nn.Linear, torch.nn.ReLU, torch.nn.BatchNorm1d)) for m in mods):
            return self.nn(h)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, torch.nn.Linear):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ReLU)
                h = F_mgcn.linear(h, m.weight, m.bias, act="relu" if relu else None, weight_layout="out_in")
                i += 2 if relu else 1
            elif isinstance(m, torch.nn.BatchNorm1d):
                h = F_mgcn.batch_norm(h, m)
                i += 1
            else:
                h = torch.relu(h)
                i += 1
        return h

    def __repr__(self):
        return "{}(nn={})".format(self.__class__.__name__, self.nn)


def global_add_pool(x, batch, size=None):
    return F_mgcn.pool_by_batch(x, batch, size, "add")


def global_mean_pool(x, batch, size=None):
    return F_mgcn.pool_by_batch(x, batch, size, "mean")


class JumpingKnowledge(torch.nn.Module):
    """'cat' and 'max' (parameter-free); 'lstm' is not used by the hot-path nets' defaults."""

    def __init__(self, mode, channels=None, num_layers=None):
        super().__init__()
        self.mode = mode.lower()
        if self.mode not in ("cat", "max"):
            raise NotImplementedError("JumpingKnowledge('lstm') is outside the hot path")

    def reset_parameters(self):
        pass

    def forward(self, xs):
        if self.mode == "cat":
            return torch.cat(xs, dim=-1)
        return torch.stack(xs, dim=-1).max(dim=-1)[0]
