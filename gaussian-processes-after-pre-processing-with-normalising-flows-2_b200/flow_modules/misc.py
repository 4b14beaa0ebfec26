"""Reduction helpers with the reference's names (flow_modules/misc.py:9-36)."""
import torch


def _as_dims(dim):
    return sorted([dim] if isinstance(dim, int) else list(dim))


def cpd_sum(tensor, dim=None, keepdim=False):
    """Sum over `dim` (int or list).  The reference reduces one dim at a time in ascending order;
    a single fused reduction gives the same value up to fp32 summation order."""
    if dim is None:
        return torch.sum(tensor)
    return tensor.sum(dim=_as_dims(dim), keepdim=keepdim)


def cpd_mean(tensor, dim=None, keepdims=False):
    if dim is None:
        return tensor.mean()
    return tensor.mean(dim=_as_dims(dim), keepdim=keepdims)
