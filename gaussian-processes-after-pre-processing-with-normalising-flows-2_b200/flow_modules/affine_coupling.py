"""Affine coupling with the reference's names and state-dict keys
(flow_modules/affine_coupling.py:10-131).  The coupling arithmetic is one fused flowk kernel;
the conditioner's ActNorm layers are folded into the convolution weights outside of init."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .common_modules import Actnormlayer, _batch_ldj


def _conv(x, weight, bias, stride, padding):
    """conv2d of the conditioner: stride-1 "same" 1x1 / 3x3 convolutions of CUDA tensors run on the flowk tcgen05 GEMMs
    (forward, input and weight gradients: flowk.tc_autograd.conv2d); anything else is a library call and needs
    FLOWK_ALLOW_LIBRARY=1 (flowk._lib.library_fallback)."""
    from .. import _lib, tc_autograd
    kh, kw = weight.shape[2], weight.shape[3]
    same = tuple(stride) == (1, 1) and tuple(padding) == (kh // 2, kw // 2) and kh == kw
    if same and tc_autograd.conv_supported(x, weight):
        return tc_autograd.conv2d(x, weight, bias, kh // 2)
    _lib.library_fallback("conv %s on input %s" % (tuple(weight.shape), tuple(x.shape)), x)
    return F.conv2d(x, weight, bias, stride, padding)


def _same_padding(kernel_size, stride):
    if isinstance(kernel_size, int):
        kernel_size = [kernel_size, kernel_size]
    if isinstance(stride, int):
        stride = [stride, stride]
    return [((k - 1) * s + 1) // 2 for k, s in zip(kernel_size, stride)]


def _resolve_padding(padding, kernel_size, stride):
    if not isinstance(padding, str):
        return padding
    mode = padding.lower()
    if mode == "same":
        return _same_padding(kernel_size, stride)
    if mode == "valid":
        return [0 for _ in (kernel_size if not isinstance(kernel_size, int) else [0, 0])]
    raise ValueError("{} is not supported".format(padding))


class Conv2dZeros(nn.Conv2d):
    """Zero-initialised conv whose output is scaled by exp(logs * logscale_factor)."""

    def __init__(self, in_channels, out_channels, kernel_size=[3, 3], stride=[1, 1], padding="same",
                 logscale_factor=3):
        super().__init__(in_channels, out_channels, kernel_size, stride,
                         _resolve_padding(padding, kernel_size, stride))
        self.logscale_factor = logscale_factor
        self.register_parameter("logs", nn.Parameter(torch.zeros(out_channels, 1, 1)))
        self.weight.data.zero_()
        self.bias.data.zero_()

    def forward(self, input):
        gain = torch.exp(self.logs.view(-1) * self.logscale_factor)
        return _conv(input, self.weight * gain.view(-1, 1, 1, 1), self.bias * gain, self.stride, self.padding)


class Conv2d(nn.Conv2d):
    """Conv (no bias) followed by ActNorm; normal(0, weight_std) init."""

    @staticmethod
    def get_padding(padding, kernel_size, stride):
        return _resolve_padding(padding, kernel_size, stride)

    def __init__(self, in_channels, out_channels, kernel_size=[3, 3], stride=[1, 1], padding="same",
                 do_actnorm=True, weight_std=0.05):
        super().__init__(in_channels, out_channels, kernel_size, stride,
                         _resolve_padding(padding, kernel_size, stride), bias=(not do_actnorm))
        self.weight.data.normal_(mean=0.0, std=weight_std)
        if not do_actnorm:
            self.bias.data.zero_()
        else:
            self.actnorm = Actnormlayer(out_channels)
        self.do_actnorm = do_actnorm

    def forward(self, input):
        if not self.do_actnorm:
            return _conv(input, self.weight, self.bias, self.stride, self.padding)
        an = self.actnorm
        if an.training and not an._seen_initialized:
            x = _conv(input, self.weight, None, self.stride, self.padding)
            x, _ = an(x, None)                     # runs the data-dependent init on the raw conv output
            return x
        # (conv(x) + b) e^{logs} == conv(x; w e^{logs}) + b e^{logs}
        gain = torch.exp(an.logs.view(-1))
        return _conv(input, self.weight * gain.view(-1, 1, 1, 1), an.bias.view(-1) * gain, self.stride, self.padding)


class NN_net(nn.Module):
    def __init__(self, in_channels, out_channels, hiddden_channels):
        super().__init__()
        self.conv1 = Conv2d(in_channels, hiddden_channels)
        self.conv2 = Conv2d(hiddden_channels, hiddden_channels, kernel_size=[1, 1])
        self.conv3 = Conv2dZeros(hiddden_channels, out_channels)

    def forward(self, x):
        """Inference (eval, autograd off) on supported shapes: three tcgen05 implicit GEMMs with the ActNorms folded
        and the ReLUs in the epilogues (flowk.conditioner_tc); otherwise the torch layers."""
        from .. import conditioner_tc
        if (conditioner_tc.ENABLED and not self.training and not torch.is_grad_enabled() and x.is_cuda
                and x.dim() == 4 and x.stride(3) == 1 and x.stride(2) == x.size(3) and x.stride(1) == x.size(2) * x.size(3)
                and conditioner_tc.affine_supported(self, x.size(2), x.size(3))):
            return conditioner_tc.affine_nn_net(self, x)
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        return self.conv3(x)


def split_feature(tensor, _type="split"):
    c = tensor.size(1)
    if _type == "split":
        return tensor[:, :c // 2, ...], tensor[:, c // 2:, ...]
    if _type == "cross":
        return tensor[:, 0::2, ...], tensor[:, 1::2, ...]
    raise ValueError(_type)


class AffineCoupling(nn.Module):
    """z = cat(z1, z2 * sigmoid(raw + 2) + shift) with (shift, raw) = NN_net(z1) de-interleaved;
    second half transformed (affine_coupling.py:100-124)."""

    def __init__(self, in_channels, out_channels, hiddden_channels):
        super().__init__()
        self.NN_net = NN_net(in_channels // 2, out_channels, hiddden_channels)

    def split(self, x, _type="split"):
        return split_feature(x, _type)

    def forward(self, input, logdet=0., reverse=False):
        ldj, had = _batch_ldj(logdet, input)
        h = self.NN_net(input[:, :input.size(1) // 2])
        z, out = ops.affine_coupling(input, h, ldj, bool(reverse))
        return z, (out if had else None)

    def forward_inference(self, x, logdet):
        return self.forward(x, logdet, reverse=False)

    def reverse_sampling(self, x, logdet):
        return self.forward(x, logdet, reverse=True)
