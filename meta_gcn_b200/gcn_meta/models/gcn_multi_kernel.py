"""Mirror of src/gcn_meta/models/gcn_multi_kernel.py: K node models over K edge sets, combined by
add / cat / mean (gcn_multi_kernel.py:76-114): the additive node model (the hot path) and the soft-attention one."""
import torch
import torch.nn as nn

from .gcn_base_models import NodeModelAdditive
from .graph_attention import NodeModelAttention


class GCNMultiKernel(nn.Module):
    nodemodel_dict = {"additive": NodeModelAdditive, "attention": NodeModelAttention}

    def __init__(self, *args, num_kernel=1, nodemodel="additive", kernel_combine="add", **kwargs):
        assert kernel_combine in ["add", "cat", "mean"]
        if nodemodel not in self.nodemodel_dict:
            raise NotImplementedError(f"nodemodel={nodemodel!r} (MLP / hard-attention node models) is outside the hot path")
        super().__init__()
        self.kernel_combine = kernel_combine
        self.node_models = nn.ModuleList(
            [self.nodemodel_dict[nodemodel](*args, **kwargs) for _ in range(num_kernel)])

    def reset_parameters(self):
        for net in self.node_models:
            net.reset_parameters()

    @staticmethod
    def _as_list(v, k):
        if isinstance(v, torch.Tensor):
            return [v]
        return [None] * k if v is None else list(v)

    def forward(self, x, edge_index_K, edge_attr_K=None, deg_K=None, edge_weight_K=None, **kwargs):
        edge_index_K = self._as_list(edge_index_K, 1)
        k = len(edge_index_K)
        per_kernel = zip(self.node_models, edge_index_K, self._as_list(edge_attr_K, k),
                         self._as_list(deg_K, k), self._as_list(edge_weight_K, k))
        act = kwargs.get("_act")
        fuse = act is not None and k == 1 and len(self.node_models) == 1
        if not fuse:
            kwargs = {kk: v for kk, v in kwargs.items() if kk != "_act"}
        outs = [nm(x, ei, ea, dg, ew, **kwargs) for nm, ei, ea, dg, ew in per_kernel if ei is not None]
        if self.kernel_combine == "cat":
            xo = torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]
        else:
            xo = outs[0]
            for o in outs[1:]:
                xo = xo + o
            if self.kernel_combine == "mean":
                xo = xo / len(outs)
        if act == "relu" and not fuse:
            xo = torch.relu(xo)
        return xo
