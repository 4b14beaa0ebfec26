"""Flow++ conditioner network with the reference's module tree and state-dict keys
(flow_modules/mixlogcdf_nn.py:8-276): weight-normalised convs, gated conv + gated 4-head
self-attention residual blocks with LayerNorm in NHWC.

Weight norm is the old `weight_g` / `weight_v` parametrisation, stated explicitly here instead of
through the deprecated torch hook, so the normalised weight can be cached in no-grad mode.
`NN.forward_raw` returns the un-split out_conv tensor [B,(2+3K)c,H,W] that the fused coupling
kernel reads in place; `NN.forward` applies the reference's post-processing for API parity.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib


def concat_elu(x, dim=1):
    from .. import tc_autograd
    if tc_autograd.pointwise_ok(x):
        return tc_autograd.concat_elu(x, dim)          # one fused kernel each way instead of neg + cat + elu
    _lib.library_fallback("concat_elu on %s" % x.dtype, x)
    return F.elu(torch.cat((x, -x), dim=dim))


def _glu(x, dim):
    from .. import tc_autograd
    if tc_autograd.pointwise_ok(x):
        return tc_autograd.glu(x, dim)
    _lib.library_fallback("GLU on %s" % x.dtype, x)
    a, b = x.chunk(2, dim=dim)
    return a * torch.sigmoid(b)


class _WeightNormed(nn.Module):
    """Holds weight_g / weight_v (/ bias) and yields w = g * v / ||v||, norm over all dims but 0."""

    def __init__(self, weight, bias):
        super().__init__()
        with torch.no_grad():
            norm = weight.reshape(weight.shape[0], -1).norm(dim=1).view(-1, *([1] * (weight.dim() - 1)))
        self.weight_g = nn.Parameter(norm.clone())
        self.weight_v = nn.Parameter(weight.detach().clone())
        if bias is None:
            self.register_parameter("bias", None)
        else:
            self.bias = nn.Parameter(bias.detach().clone())
        self._cache = None

    def normed_weight(self):
        v, g = self.weight_v, self.weight_g
        track = torch.is_grad_enabled() and (v.requires_grad or g.requires_grad)
        if not track:
            key = _lib.param_key((v, g))
            if self._cache is not None and self._cache[0] == key:
                return self._cache[1]
        norm = v.reshape(v.shape[0], -1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
        w = v * (g / norm)
        if not track:
            self._cache = (key, w)
        return w


class _WNConvCore(_WeightNormed):
    def __init__(self, in_channels, out_channels, kernel_size, padding, bias):
        ref = nn.Conv2d(in_channels, out_channels, kernel_size, padding=padding, bias=bias)   # default init
        super().__init__(ref.weight, ref.bias)
        self.padding = padding

    def forward(self, x):
        from .. import tc_autograd
        v = self.weight_v
        prepared = tc_autograd.take_prepared(self)      # always consumed: operands are valid for ONE forward after a refresh
        if self.padding == v.shape[2] // 2 and tc_autograd.conv_supported(x, v):
            # weight norm + operands fused, forward (+ dgrad, wgrad under autograd) on tcgen05
            return tc_autograd.wn_conv2d(x, v, self.weight_g, self.bias, prepared)
        _lib.library_fallback("weight-normed conv %s on input %s" % (tuple(v.shape), tuple(x.shape)), x)
        return F.conv2d(x, self.normed_weight(), self.bias, padding=self.padding)


class _WNLinear(_WeightNormed):
    def __init__(self, in_features, out_features, bias=True):
        ref = nn.Linear(in_features, out_features, bias=bias)
        super().__init__(ref.weight, ref.bias)

    def forward(self, x):
        from .. import tc_autograd
        prepared = tc_autograd.take_prepared(self)
        if tc_autograd.linear_supported(x, self.weight_v):
            return tc_autograd.wn_linear(x, self.weight_v, self.weight_g, self.bias, prepared)
        _lib.library_fallback("weight-normed linear %s on input %s" % (tuple(self.weight_v.shape), tuple(x.shape)), x)
        return F.linear(x, self.normed_weight(), self.bias)


class WNConv2d(nn.Module):
    """Weight-normalised conv; parameters live under `.conv.` as in the reference (mixlogcdf_nn.py:12-29)."""

    def __init__(self, in_channels, out_channels, kernel_size, padding, bias=True):
        super().__init__()
        self.conv = _WNConvCore(in_channels, out_channels, kernel_size, padding, bias)

    def forward(self, x):
        return self.conv(x)


class GatedConv(nn.Module):
    """concat_elu -> WN conv3x3 (2C->C) -> concat_elu -> Dropout2d -> WN conv1x1 (2C->2C) -> GLU."""

    def __init__(self, num_channels, drop_prob=0., aux_channels=None):
        super().__init__()
        self.nlin = concat_elu
        self.conv = WNConv2d(2 * num_channels, num_channels, kernel_size=3, padding=1)
        self.drop = nn.Dropout2d(drop_prob)
        self.gate = WNConv2d(2 * num_channels, 2 * num_channels, kernel_size=1, padding=0)
        self.aux_conv = (WNConv2d(2 * aux_channels, num_channels, kernel_size=1, padding=0)
                         if aux_channels is not None else None)

    def forward(self, x, aux=None):
        x = self.conv(self.nlin(x))
        if aux is not None:
            x = x + self.aux_conv(self.nlin(aux))
        from .. import tc_autograd
        if tc_autograd.pointwise_ok(x) and torch.is_grad_enabled() and tc_autograd.FOLD_DROPOUT:
            # concat_elu with the feature dropout that follows it folded in (one kernel each way)
            mask = None
            if self.training and self.drop.p > 0:
                mask = tc_autograd.feature_dropout_mask(x, 2 * x.shape[1], self.drop.p)
            x = self.gate(tc_autograd.concat_elu(x, 1, mask))
        else:
            x = self.gate(self.drop(self.nlin(x)))
        return _glu(x, 1)


class GatedAttn(nn.Module):
    """Gated multi-head self-attention over the H*W positions of an NHWC tensor (mixlogcdf_nn.py:105-224)."""

    def __init__(self, d_model, num_heads=4, drop_prob=0.):
        super().__init__()
        self.d_model = d_model
        self.num_heads = num_heads
        self.drop_prob = drop_prob
        self.in_proj = _WNLinear(d_model, 3 * d_model, bias=False)
        self.gate = _WNLinear(d_model, 2 * d_model)
        self._pos = {}
        self._salt = 1                             # keys this layer's dropout masks (flowk training attention kernels):
                                                   # its index inside the owning NN / model (see assign_dropout_salts)

    @staticmethod
    def get_pos_enc(seq_len, num_channels, device):
        half = num_channels // 2
        inc = math.log(10000.) / (half - 1)
        inv = torch.exp(torch.arange(half, dtype=torch.float32, device=device) * -inc)
        t = torch.arange(seq_len, dtype=torch.float32, device=device).unsqueeze(1) * inv.unsqueeze(0)
        enc = torch.cat([t.sin(), t.cos()], dim=1)
        enc = F.pad(enc, [0, num_channels % 2, 0, 0])
        return enc.view(1, seq_len, num_channels)

    def _pos_enc(self, seq_len, ch, device):
        key = (seq_len, ch, str(device))
        if key not in self._pos:
            self._pos[key] = self.get_pos_enc(seq_len, ch, device)
        return self._pos[key]

    def forward(self, x):
        b, h, w, c = x.shape
        seq, heads, d = h * w, self.num_heads, c // self.num_heads
        t = x.reshape(b, seq, c) + self._pos_enc(seq, c, x.device)
        proj = self.in_proj(t)
        from .. import tc_autograd
        if tc_autograd.attention_train_supported(proj, heads):
            # fused forward / backward with in-kernel dropout: nothing of size seq x seq touches HBM
            att = tc_autograd.attention_core(proj, heads, self.drop_prob if self.training else 0.0, self._salt)
            return _glu(self.gate(att.view(b, h, w, c)), -1)
        # in_proj output order is (k | v | q): memory = first 2C, query = last C (mixlogcdf_nn.py:136-139).
        # ONE strided copy puts all three head-first, [3, B, heads, seq, d] contiguous: the batched matmuls then need
        # no further layout copies and the backward is a single permute-copy instead of slice-gradient fills.
        _lib.library_fallback("attention core (seq %d, head dim %d)" % (seq, d), x)
        kvq = proj.view(b, seq, 3, heads, d).permute(2, 0, 3, 1, 4).contiguous()
        k, v, q = kvq[0], kvq[1], kvq[2] * (d ** -0.5)
        weights = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
        weights = F.dropout(weights, self.drop_prob, self.training)
        att = (weights @ v).permute(0, 2, 1, 3).reshape(b, h, w, c)
        return _glu(self.gate(att), -1)


def assign_dropout_salts(root):
    """Number the GatedAttn layers of `root` 1..n in module order: a layer's dropout stream depends on its position in
    the model, not on how many other models the process built before."""
    for i, m in enumerate(m for m in root.modules() if isinstance(m, GatedAttn)):
        m._salt = i + 1


class ConvAttnBlock(nn.Module):
    def __init__(self, num_channels, drop_prob, use_attn, aux_channels):
        super().__init__()
        self.conv = GatedConv(num_channels, drop_prob, aux_channels)
        self.norm_1 = nn.LayerNorm(num_channels)
        if use_attn:
            self.attn = GatedAttn(num_channels, drop_prob=drop_prob)
            self.norm_2 = nn.LayerNorm(num_channels)
        else:
            self.attn = None

    def forward(self, x, aux=None):
        from .. import tc_autograd
        if tc_autograd.pointwise_ok(x) and x.size(1) <= 512:
            # residual + LayerNorm (+ both permutes) as one kernel each way
            x = tc_autograd.add_layernorm(self.conv(x, aux), x, self.norm_1, True, self.attn is None)
            if self.attn:
                x = tc_autograd.add_layernorm(self.attn(x), x, self.norm_2, False, True)
            return x
        _lib.library_fallback("residual + LayerNorm over %d channels" % x.size(1), x)
        x = self.conv(x, aux) + x
        x = self.norm_1(x.permute(0, 2, 3, 1))
        if self.attn:
            x = self.norm_2(self.attn(x) + x)
        return x.permute(0, 3, 1, 2)


class Rescale(nn.Module):
    """Per-channel multiplier, wrapped in weight norm by NN (mixlogcdf_nn.py:63,263-276)."""

    def __init__(self, num_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels, 1, 1))

    def forward(self, x):
        return self.weight * x


class _WNRescale(_WeightNormed):
    """weight_norm(Rescale(c)): keys `rescale.weight_g`, `rescale.weight_v`, both [c,1,1]."""

    def __init__(self, num_channels):
        super().__init__(torch.ones(num_channels, 1, 1), None)

    def forward(self, x):
        return self.normed_weight() * x


class NN(nn.Module):
    """Conditioner of the MixLogCDF coupling (mixlogcdf_nn.py:32-78)."""

    def __init__(self, in_channels, num_channels, num_blocks, num_components, drop_prob, use_attn=True,
                 aux_channels=None):
        super().__init__()
        self.k = num_components
        self.in_conv = WNConv2d(in_channels, num_channels, kernel_size=3, padding=1)
        self.mid_convs = nn.ModuleList([ConvAttnBlock(num_channels, drop_prob, use_attn, aux_channels)
                                        for _ in range(num_blocks)])
        self.out_conv = WNConv2d(num_channels, in_channels * (2 + 3 * self.k), kernel_size=3, padding=1)
        self.rescale = _WNRescale(in_channels)
        assign_dropout_salts(self)

    def forward_raw(self, x, aux=None):
        """Un-split out_conv output [B,(2+3K)c,H,W].  In inference mode (eval, autograd off) on supported shapes
        the whole stack runs as tcgen05 implicit GEMMs (flowk.conditioner_tc); otherwise (training: dropout and
        autograd; odd channel counts) through the torch layers below."""
        from .. import conditioner_tc
        if (conditioner_tc.ENABLED and aux is None and not self.training and not torch.is_grad_enabled()
                and x.is_cuda and conditioner_tc.supported(self.in_conv.conv.weight_v.shape[0], x.size(2), x.size(3))):
            return conditioner_tc.mixlogcdf_nn_raw(self, x)
        x = self.in_conv(x)
        for block in self.mid_convs:
            x = block(x, aux)
        return self.out_conv(x)

    def rescale_weight(self):
        return self.rescale.normed_weight().reshape(-1)

    def forward(self, x, aux=None):
        b, c, h, w = x.size()
        raw = self.forward_raw(x, aux).view(b, -1, c, h, w)
        s, t, pi, mu, scales = raw.split((1, 1, self.k, self.k, self.k), dim=1)
        s = self.rescale(torch.tanh(s.squeeze(1)))
        return s, t.squeeze(1), pi, mu, scales.clamp(min=-7)
