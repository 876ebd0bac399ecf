"""kernel/graph_sage.py mirror: GraphSAGE / GraphSAGEWithJK = num_layers x relu(SAGEConv) -> mean
pool -> MLP head (kernel/graph_sage.py:7-36, 39-77).  ReLU is fused into the transform epilogue."""
import torch

from ..compat.torch_geometric.nn import SAGEConv
from ._base import GraphClassifier


class GraphSAGE(GraphClassifier):
    def __init__(self, dataset, num_layers, hidden):
        super().__init__()
        self.conv1 = SAGEConv(dataset.num_features, hidden)
        self.convs = torch.nn.ModuleList(SAGEConv(hidden, hidden) for _ in range(num_layers - 1))
        self._init_head(dataset, num_layers, hidden)

    def _conv(self, conv, x, edge_index):
        return conv(x, edge_index, _act="relu")


class GraphSAGEWithJK(GraphSAGE):
    def __init__(self, dataset, num_layers, hidden, mode="cat"):
        torch.nn.Module.__init__(self)
        self.conv1 = SAGEConv(dataset.num_features, hidden)
        self.convs = torch.nn.ModuleList(SAGEConv(hidden, hidden) for _ in range(num_layers - 1))
        self._init_head(dataset, num_layers, hidden, mode)
