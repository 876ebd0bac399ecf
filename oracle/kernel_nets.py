"""ORACLE / TEST INFRASTRUCTURE ONLY — CPU restatement of the kernel/ benchmark nets on the hot path
(kernel/gcn.py:7-77, kernel/gin.py:7-160, kernel/graph_sage.py:7-77) over the PyG-1.3 operator
restatement in oracle/shim.  Attribute names match the reference (conv1, convs, jump, lin1, lin2) so
state_dicts move between the reference, this oracle and meta_gcn_b200.kernel.  Checked against the
reference's own files by tests/test_oracle_golden.py (goldens from oracle/make_golden.py)."""
import torch
import torch.nn.functional as F
from torch.nn import BatchNorm1d, Linear, ReLU, Sequential

from . import use_shim

use_shim()
from torch_geometric.nn import GCNConv, GINConv, JumpingKnowledge, SAGEConv, global_mean_pool  # noqa: E402


def _conv(kind, cin, hidden):
    if kind == "gcn":
        return GCNConv(cin, hidden)                                      # kernel/gcn.py:10,13
    if kind == "sage":
        return SAGEConv(cin, hidden)                                     # kernel/graph_sage.py:10,13
    mlp = Sequential(Linear(cin, hidden), ReLU(), Linear(hidden, hidden), ReLU(), BatchNorm1d(hidden))
    return GINConv(mlp, train_eps=(kind == "gin"))                       # kernel/gin.py:10-17,119-127


class OracleGraphNet(torch.nn.Module):
    """kind: 'gcn' | 'sage' | 'gin0' | 'gin'; jk: None | 'cat' | 'max'."""

    def __init__(self, kind, num_features, num_classes, num_layers, hidden, jk=None, dropout=True):
        super().__init__()
        self.kind, self.use_dropout = kind, dropout
        self.conv1 = _conv(kind, num_features, hidden)
        self.convs = torch.nn.ModuleList(_conv(kind, hidden, hidden) for _ in range(num_layers - 1))
        if jk is not None:
            self.jump = JumpingKnowledge(jk)
        self.lin1 = Linear(num_layers * hidden if jk == "cat" else hidden, hidden)
        self.lin2 = Linear(hidden, num_classes)

    def forward(self, data):
        x, edge_index, batch = data.x, data.edge_index, data.batch
        relu_after = self.kind in ("gcn", "sage")                        # gin.py:41-43 has no outer relu
        xs = []
        for conv in [self.conv1, *self.convs]:
            x = conv(x, edge_index)
            x = F.relu(x) if relu_after else x
            xs.append(x)
        if hasattr(self, "jump"):
            x = self.jump(xs)
        x = global_mean_pool(x, batch)
        x = F.relu(self.lin1(x))
        if self.use_dropout:
            x = F.dropout(x, p=0.5, training=self.training)
        x = self.lin2(x)
        return F.log_softmax(x, dim=-1)
