"""Mixture-of-logistics CDF coupling (flow_modules/mixlogcdf_coupling.py:17-57): conditioner NN on
the second half, fused flowk kernel for everything after out_conv - parameter split, tanh/rescale,
scale clamp, mixture CDF, logit, affine, log-det reduction, concatenation (and, from FlowStep, the
TupleFlip that follows)."""
import torch.nn as nn

from .. import ops
from .common_modules import _batch_ldj
from .mixlogcdf_nn import NN


def split_feature(tensor, _type="split"):
    c = tensor.size(1)
    if _type == "split":
        return tensor[:, :c // 2, ...], tensor[:, c // 2:, ...]
    if _type == "cross":
        return tensor[:, 0::2, ...], tensor[:, 1::2, ...]
    raise ValueError(_type)


class MixLogCDFCoupling(nn.Module):
    def __init__(self, in_channels, mid_channels, num_blocks, num_components, drop_prob, use_attn=True,
                 aux_channels=None):
        super().__init__()
        self.nn = NN(in_channels // 2, mid_channels, num_blocks, num_components, drop_prob, use_attn, aux_channels)
        self.num_components = num_components

    def split(self, x, _type="split"):
        return split_feature(x, _type)

    def forward(self, x, sldj=None, reverse=False, aux=None, flip=False):
        """`flip=True` fuses the TupleFlip that FlowStep applies after (forward) / before (reverse)
        this layer: forward returns cat(x_id, out); reverse expects cat(x_id, v) and returns cat(x, x_id)."""
        c = x.size(1) // 2
        x_id = x[:, :c] if (reverse and flip) else x[:, c:]
        ldj, had = _batch_ldj(sldj, x)
        raw = self.nn.forward_raw(x_id, aux)
        y, out = ops.mixlogcdf_coupling(x, raw, self.nn.rescale_weight(), ldj, bool(reverse), bool(flip),
                                        self.num_components)
        return y, (out if had else None)
