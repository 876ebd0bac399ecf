"""Mirror of src/gcn_meta/models/common.py for the hot path: the ``scatter_`` primitive seam
(common.py:37-66) and the activation factory (common.py:27-34), backed by libmgcn kernels."""
import torch
import torch.nn as nn

from ... import functional as F_mgcn


class Identity(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, input):
        return input


_ACTIVATIONS = {
    "lrelu": lambda slope: nn.LeakyReLU(slope),
    "relu": lambda slope: nn.ReLU(),
    "elu": lambda slope: nn.ELU(),
    "none": lambda slope: Identity(),
}


def activation(act, negative_slope=0.2):
    return _ACTIVATIONS[act](negative_slope)


def scatter_(name, src, index, dim_size=None, out=None):
    """Row-wise aggregation of ``src`` by ``index`` along dim 0 ('add' | 'mean' | 'max').

    Same call shape as the reference (common.py:37); computed by the row-owned gather-sum kernel
    over a stable sort of ``index`` — deterministic, and for rows below the hub threshold the same
    fp32 summation order as the reference's CPU scatter_add.  'max' (common.py:57,63-64: fill -1e38, untouched
    rows set to 0) is a row-owned first-maximum pass with torch_scatter's gradient rule."""
    assert name in ["add", "mean", "max"]
    if name == "max":
        res = F_mgcn.scatter_rows_max(src, index, dim_size)[0]
        if out is not None:
            res = torch.maximum(out, res)
        return res
    res = F_mgcn.scatter_rows(src, index, dim_size, name)
    if out is not None:
        res = out + res if name == "add" else res
    return res


def softmax(src, index, num_nodes=None):
    """Sparsely evaluated softmax over the groups of ``index`` (common.py:69-92): segment max for stability
    (fill -1e16 as in the reference), exp, segment sum + 1e-16 — on the libmgcn scatter kernels"""
    from ...compat.torch_scatter import scatter_add, scatter_max
    if num_nodes is None:
        num_nodes = int(index.max().item()) + 1
    out = src - scatter_max(src, index, dim=0, dim_size=num_nodes, fill_value=-1e16)[0][index]
    out = out.exp()
    out = out / (scatter_add(out, index, dim=0, dim_size=num_nodes)[index] + 1e-16)
    return out
