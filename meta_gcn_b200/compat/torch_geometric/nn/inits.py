"""Initialisers the reference imports from torch_geometric.nn.inits (gcn_base_models.py:5)."""
import math


def uniform(size, tensor):
    if tensor is not None:
        bound = 1.0 / math.sqrt(size)
        tensor.data.uniform_(-bound, bound)


def glorot(tensor):
    if tensor is not None:
        bound = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-bound, bound)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)


def reset(module):
    if module is None:
        return
    children = list(module.children()) if hasattr(module, "children") else []
    for item in (children if children else [module]):
        if hasattr(item, "reset_parameters"):
            item.reset_parameters()
