"""kernel/gcn.py mirror: GCN / GCNWithJK = num_layers x relu(GCNConv) -> mean pool -> MLP head
(kernel/gcn.py:7-36, 39-77).  ReLU is fused into the aggregation epilogue."""
import torch

from ..compat.torch_geometric.nn import GCNConv
from ._base import GraphClassifier


class GCN(GraphClassifier):
    def __init__(self, dataset, num_layers, hidden):
        super().__init__()
        self.conv1 = GCNConv(dataset.num_features, hidden)
        self.convs = torch.nn.ModuleList(GCNConv(hidden, hidden) for _ in range(num_layers - 1))
        self._init_head(dataset, num_layers, hidden)

    def _conv(self, conv, x, edge_index):
        return conv(x, edge_index, _act="relu")


class GCNWithJK(GCN):
    def __init__(self, dataset, num_layers, hidden, mode="cat"):
        torch.nn.Module.__init__(self)
        self.conv1 = GCNConv(dataset.num_features, hidden)
        self.convs = torch.nn.ModuleList(GCNConv(hidden, hidden) for _ in range(num_layers - 1))
        self._init_head(dataset, num_layers, hidden, mode)
