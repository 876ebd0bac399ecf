"""kernel/gin.py mirror: GIN0 (eps fixed) / GIN (eps learnt) and their JK variants =
num_layers x GINConv(Linear-ReLU-Linear-ReLU-BatchNorm1d) -> mean pool -> MLP head
(kernel/gin.py:7-48, 117-160)."""
import torch
from torch.nn import BatchNorm1d as BN
from torch.nn import Linear, ReLU, Sequential

from ..compat.torch_geometric.nn import GINConv
from ._base import GraphClassifier


def _gin_mlp(cin, hidden):
    return Sequential(Linear(cin, hidden), ReLU(), Linear(hidden, hidden), ReLU(), BN(hidden))


class _GINBase(GraphClassifier):
    train_eps = False

    def __init__(self, dataset, num_layers, hidden, mode=None):
        super().__init__()
        self.conv1 = GINConv(_gin_mlp(dataset.num_features, hidden), train_eps=self.train_eps)
        self.convs = torch.nn.ModuleList(
            GINConv(_gin_mlp(hidden, hidden), train_eps=self.train_eps) for _ in range(num_layers - 1))
        self._init_head(dataset, num_layers, hidden, mode)

    def _conv(self, conv, x, edge_index):
        return conv(x, edge_index)


class GIN0(_GINBase):
    def __init__(self, dataset, num_layers, hidden):
        super().__init__(dataset, num_layers, hidden)


class GIN0WithJK(_GINBase):
    def __init__(self, dataset, num_layers, hidden, mode="cat"):
        super().__init__(dataset, num_layers, hidden, mode)


class GIN(_GINBase):
    train_eps = True

    def __init__(self, dataset, num_layers, hidden):
        super().__init__(dataset, num_layers, hidden)


class GINWithJK(_GINBase):
    train_eps = True

    def __init__(self, dataset, num_layers, hidden, mode="cat"):
        super().__init__(dataset, num_layers, hidden, mode)
