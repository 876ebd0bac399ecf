"""Shared skeleton of the kernel/ nets: conv stack -> (JumpingKnowledge) -> global_mean_pool ->
relu(lin1) -> dropout(0.5) -> lin2 -> log_softmax (kernel/gcn.py:24-33, gin.py:39-48,
graph_sage.py:24-33).  Attribute names (conv1, convs, jump, lin1, lin2) follow the reference so
state_dicts are interchangeable."""
import torch
import torch.nn.functional as F
from torch.nn import Linear

from .. import functional as F_mgcn
from ..compat.torch_geometric.nn import JumpingKnowledge, global_mean_pool


class GraphClassifier(torch.nn.Module):
    def _init_head(self, dataset, num_layers, hidden, mode=None):
        if mode is not None:
            self.jump = JumpingKnowledge(mode)
        width = num_layers * hidden if mode == "cat" else hidden
        self.lin1 = Linear(width, hidden)
        self.lin2 = Linear(hidden, dataset.num_classes)

    def _convs(self):
        return [self.conv1, *self.convs]

    def reset_parameters(self):
        for conv in self._convs():
            conv.reset_parameters()
        if hasattr(self, "jump"):
            self.jump.reset_parameters()
        self.lin1.reset_parameters()
        self.lin2.reset_parameters()

    def _conv(self, conv, x, edge_index):
        raise NotImplementedError

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        xs = []
        for conv in self._convs():
            x = self._conv(conv, x, edge_index)
            xs.append(x)
        if hasattr(self, "jump"):
            x = self.jump(xs)
        size = getattr(data, "num_graphs", None)
        x = global_mean_pool(x, batch, size)
        x = F_mgcn.linear(x, self.lin1.weight, self.lin1.bias, act="relu", weight_layout="out_in")
        x = F.dropout(x, p=0.5, training=self.training)
        x = F_mgcn.linear(x, self.lin2.weight, self.lin2.bias, weight_layout="out_in")
        return F.log_softmax(x, dim=-1)

    def __repr__(self):
        return self.__class__.__name__
