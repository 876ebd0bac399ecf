"""ORACLE ONLY — torch_sparse.utils.unique (torch_sparse 0.4.x): sorted unique values and, for each,
the position of one of its occurrences in the input."""
import torch


def unique(src):
    src = src.contiguous().view(-1)
    output, inverse = torch.unique(src, sorted=True, return_inverse=True)
    perm = torch.arange(inverse.size(0), dtype=inverse.dtype, device=inverse.device)
    perm = inverse.new_empty(output.size(0)).scatter_(0, inverse, perm)
    return output, perm
