"""CPU oracle for the mAR-SCF flow-step hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functions over a flat ``{key: tensor}`` state
dict, what the reference's ``nn.Module`` stack computes on the hot path that
BASELINE.json's ``north_star`` names.  It is the checker
for the CUDA kernels; nothing in the product package may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs are allowed to.

Pinning: ``tests/golden/*.npz`` hold inputs/outputs produced by the reference's
own modules (imported read-only from /root/reference by
``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks every
function below against them.  The reference itself ships no known-answer
vectors, so the fixtures generated from its live modules
are the pin.

All functions are dtype-generic: float32 reproduces the reference's own
arithmetic (same ATen calls in the same order wherever rounding could matter),
float64 gives a higher-precision yardstick for tolerance budgeting.

Citations ``file:line`` are relative to the reference repository root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

LOG_FLOOR = 1e-22          # flow_modules/log_dist.py:5-6  (safe_log clamp)
U_CLAMP = 1e-5             # flow_modules/mixlogcdf_coupling.py:44
S_FLOOR = -7.0             # flow_modules/mixlogcdf_nn.py:76
BISECT_EPS = 1e-10         # flow_modules/log_dist.py:43
BISECT_MAX_ITERS = 100     # flow_modules/log_dist.py:44


# --------------------------------------------------------------------------
# reductions (flow_modules/misc.py:9-36)
# --------------------------------------------------------------------------
def seq_sum(t: Tensor, dims: Sequence[int]) -> Tensor:
    """cpd_sum: reduce the listed dims one at a time, ascending (misc.py:15-21)."""
    for d in sorted(dims):
        t = t.sum(dim=d, keepdim=True)
    for i, d in enumerate(sorted(dims)):
        t = t.squeeze(d - i)
    return t


def seq_mean_keep(t: Tensor, dims: Sequence[int]) -> Tensor:
    """cpd_mean(..., keepdims=True) (misc.py:23-36)."""
    for d in sorted(dims):
        t = t.mean(dim=d, keepdim=True)
    return t


# --------------------------------------------------------------------------
# squeeze / unsqueeze (flow_modules/common_modules.py:12-42)
# --------------------------------------------------------------------------
def squeeze2d(x: Tensor, factor: int = 2) -> Tensor:
    """out[b, c*f*f + fh*f + fw, h, w] = in[b, c, h*f + fh, w*f + fw]  (common_modules.py:22-24)."""
    if factor == 1:
        return x
    b, c, h, w = x.shape
    if h % factor or w % factor:
        raise AssertionError("{}".format((h, w)))          # common_modules.py:21
    out = x.new_empty(b, c * factor * factor, h // factor, w // factor)
    for fh in range(factor):
        for fw in range(factor):
            out[:, (fh * factor + fw)::factor * factor] = x[:, :, fh::factor, fw::factor]
    return out


def unsqueeze2d(x: Tensor, factor: int = 2) -> Tensor:
    """Exact inverse permutation of :func:`squeeze2d` (common_modules.py:39-41)."""
    if factor == 1:
        return x
    b, c, h, w = x.shape
    f2 = factor * factor
    if c % f2:
        raise AssertionError("{}".format(c))               # common_modules.py:38
    out = x.new_empty(b, c // f2, h * factor, w * factor)
    for fh in range(factor):
        for fw in range(factor):
            out[:, :, fh::factor, fw::factor] = x[:, (fh * factor + fw)::f2]
    return out


# --------------------------------------------------------------------------
# ActNorm (flow_modules/common_modules.py:130-186)
# --------------------------------------------------------------------------
def actnorm_init(x: Tensor, scale: float = 1.0, eps: float = 1e-6) -> Tuple[Tensor, Tensor]:
    """Data-dependent init (common_modules.py:141-151): returns (bias, logs) shaped [1,C,1,1]."""
    bias = -seq_mean_keep(x, [0, 2, 3])
    var = seq_mean_keep((x + bias) ** 2, [0, 2, 3])
    logs = (scale / (var.sqrt() + eps)).log()
    return bias, logs


def actnorm(x: Tensor, bias: Tensor, logs: Tensor, ldj: Optional[Tensor], reverse: bool = False):
    """fwd y=(x+bias)*exp(logs); rev x=y*exp(-logs)-bias; ldj +/- sum(logs)*H*W (common_modules.py:153-186)."""
    hw = x.shape[2] * x.shape[3]
    if reverse:
        y = x * torch.exp(-logs) - bias
    else:
        y = (x + bias) * torch.exp(logs)
    if ldj is not None:
        d = logs.sum() * hw
        ldj = ldj - d if reverse else ldj + d
    return y, ldj


# --------------------------------------------------------------------------
# invertible 1x1 convolution, LU parametrisation (common_modules.py:57-127)
# --------------------------------------------------------------------------
def invconv_weight(p: Tensor, l: Tensor, u: Tensor, sign_s: Tensor, log_s: Tensor,
                   reverse: bool = False) -> Tensor:
    """W = P (L.tril(-1)+I) (U.triu(1)+diag(sign_s e^{log_s}))  (common_modules.py:102-106).

    Reverse goes through float64 inverses of the two triangular factors and a
    float32 inverse of P, multiplied as U^-1 (L^-1 P^-1) (common_modules.py:108-110);
    the reference's trailing ``.cuda()`` is the only thing not restated.
    """
    c = l.shape[0]
    lower_mask = torch.tril(torch.ones(c, c, dtype=l.dtype), -1)
    eye = torch.eye(c, dtype=l.dtype)
    lo = l * lower_mask + eye
    up = u * lower_mask.t().contiguous() + torch.diag(sign_s * torch.exp(log_s))
    if not reverse:
        return p @ (lo @ up)
    lo_inv = torch.inverse(lo.double()).to(l.dtype)
    up_inv = torch.inverse(up.double()).to(l.dtype)
    return up_inv @ (lo_inv @ torch.inverse(p))


def invconv(x: Tensor, p: Tensor, l: Tensor, u: Tensor, sign_s: Tensor, log_s: Tensor,
            ldj: Optional[Tensor], reverse: bool = False):
    """z[b,:,h,w] = W x[b,:,h,w]; ldj +/- sum(log_s) * W_spatial**2 (common_modules.py:86,104,113-127).

    The log-det multiplier is the LAST spatial dim squared, as in the reference
    (correct for square maps only; reproduced, not fixed).
    """
    w = invconv_weight(p, l, u, sign_s, log_s, reverse)
    px = x.shape[-1]
    d = log_s.sum() * px * px
    z = F.conv2d(x, w.view(w.shape[0], w.shape[1], 1, 1))
    if ldj is not None:
        ldj = ldj - d if reverse else ldj + d
    return z, ldj


# --------------------------------------------------------------------------
# affine coupling (flow_modules/affine_coupling.py)
# --------------------------------------------------------------------------
def affine_conditioner(sd: State, pre: str, z1: Tensor) -> Tensor:
    """NN_net (affine_coupling.py:68-80): conv3x3 -> actnorm -> relu -> conv1x1 -> actnorm -> relu
    -> zero-init conv3x3 (+bias) scaled by exp(3*logs) (affine_coupling.py:23-25)."""
    h = F.conv2d(z1, sd[pre + "conv1.weight"], None, padding=1)
    h, _ = actnorm(h, sd[pre + "conv1.actnorm.bias"], sd[pre + "conv1.actnorm.logs"], None)
    h = F.relu(h)
    h = F.conv2d(h, sd[pre + "conv2.weight"], None, padding=0)
    h, _ = actnorm(h, sd[pre + "conv2.actnorm.bias"], sd[pre + "conv2.actnorm.logs"], None)
    h = F.relu(h)
    h = F.conv2d(h, sd[pre + "conv3.weight"], sd[pre + "conv3.bias"], padding=1)
    return h * torch.exp(sd[pre + "conv3.logs"] * 3.0)


def affine_elementwise(x: Tensor, h: Tensor, ldj: Tensor, reverse: bool = False):
    """The coupling arithmetic alone, given the conditioner output h (affine_coupling.py:103-124).

    second half transformed; shift = h[:,0::2], scale = sigmoid(h[:,1::2] + 2).
    """
    c = x.shape[1] // 2
    z1, z2 = x[:, :c], x[:, c:]
    shift, raw = h[:, 0::2], h[:, 1::2]
    scale = torch.sigmoid(raw + 2.0)
    d = seq_sum(torch.log(scale), [1, 2, 3])
    if not reverse:
        z2 = z2 * scale
        z2 = shift + z2
        ldj = d + ldj
    else:
        z2 = z2 - shift
        z2 = z2 / scale
        ldj = ldj - d
    return torch.cat((z1, z2), dim=1), ldj


def affine_coupling(sd: State, pre: str, x: Tensor, ldj: Tensor, reverse: bool = False):
    """AffineCoupling.forward (affine_coupling.py:126-131); ``pre`` ends in 'coupling.'."""
    c = x.shape[1] // 2
    h = affine_conditioner(sd, pre + "NN_net.", x[:, :c])
    return affine_elementwise(x, h, ldj, reverse)


# --------------------------------------------------------------------------
# Flow++ conditioner (flow_modules/mixlogcdf_nn.py)
# --------------------------------------------------------------------------
def wn(sd: State, pre: str) -> Tensor:
    """Old-style weight_norm, dim=0: w = g * v / ||v||_(all dims but 0) (mixlogcdf_nn.py:23-24,63,121-122)."""
    v, g = sd[pre + "weight_v"], sd[pre + "weight_g"]
    norm = v.reshape(v.shape[0], -1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def concat_elu(x: Tensor, dim: int = 1) -> Tensor:
    """mixlogcdf_nn.py:8-10."""
    return F.elu(torch.cat((x, -x), dim=dim))


def positional_encoding(seq_len: int, ch: int, dtype) -> Tensor:
    """Sinusoid table (mixlogcdf_nn.py:208-224), [1, seq, ch]: first half sin, second half cos."""
    half = ch // 2
    inc = math.log(10000.0) / (half - 1)
    inv = torch.exp(torch.arange(half, dtype=torch.float32) * -inc)
    t = torch.arange(seq_len, dtype=torch.float32).unsqueeze(1) * inv.unsqueeze(0)
    enc = torch.cat([t.sin(), t.cos()], dim=1)
    enc = F.pad(enc, [0, ch % 2, 0, 0])
    return enc.view(1, seq_len, ch).to(dtype)


def gated_conv(sd: State, pre: str, x: Tensor) -> Tensor:
    """GatedConv.forward, eval mode (dropout off) (mixlogcdf_nn.py:248-260)."""
    h = concat_elu(x)
    h = F.conv2d(h, wn(sd, pre + "conv.conv."), sd[pre + "conv.conv.bias"], padding=1)
    h = concat_elu(h)
    h = F.conv2d(h, wn(sd, pre + "gate.conv."), sd[pre + "gate.conv.bias"], padding=0)
    a, b = h.chunk(2, dim=1)
    return a * torch.sigmoid(b)


def gated_attn(sd: State, pre: str, x: Tensor, heads: int = 4) -> Tensor:
    """GatedAttn.forward on NHWC input, eval mode (mixlogcdf_nn.py:124-152).

    in_proj output splits as (memory = first 2C -> k, v ; query = last C) (:136-139);
    q scaled by (C/heads)^-1/2 (:143-144); the head-merged result is re-read through a
    transpose(1,2).view(b,c,h,w).permute(0,2,3,1) round trip (:147) that is the identity
    on [b, h*w, c]; then a weight-normed Linear C->2C and a GLU (:149-151).
    """
    b, hh, ww, c = x.shape
    seq = hh * ww
    t = x.reshape(b, seq, c) + positional_encoding(seq, c, x.dtype)
    proj = F.linear(t, wn(sd, pre + "in_proj."))
    memory, query = proj[..., :2 * c], proj[..., 2 * c:]
    k, v = memory[..., :c], memory[..., c:]
    d = c // heads

    def split_heads(m):
        return m.reshape(b, seq, heads, d).permute(0, 2, 1, 3)

    q = split_heads(query) * (d ** -0.5)
    att = torch.softmax(q @ split_heads(k).transpose(-1, -2), dim=-1) @ split_heads(v)
    att = att.permute(0, 2, 1, 3).reshape(b, seq, c).reshape(b, hh, ww, c)
    g = F.linear(att, wn(sd, pre + "gate."), sd[pre + "gate.bias"])
    a, gate = g.chunk(2, dim=-1)
    return a * torch.sigmoid(gate)


def conv_attn_block(sd: State, pre: str, x: Tensor) -> Tensor:
    """ConvAttnBlock.forward (mixlogcdf_nn.py:92-102): NCHW in, NCHW out."""
    c = x.shape[1]
    x = gated_conv(sd, pre + "conv.", x) + x
    x = x.permute(0, 2, 3, 1)
    x = F.layer_norm(x, (c,), sd[pre + "norm_1.weight"], sd[pre + "norm_1.bias"])
    x = gated_attn(sd, pre + "attn.", x) + x
    x = F.layer_norm(x, (c,), sd[pre + "norm_2.weight"], sd[pre + "norm_2.bias"])
    return x.permute(0, 3, 1, 2)


def mixlogcdf_conditioner_raw(sd: State, pre: str, x_id: Tensor) -> Tensor:
    """NN.forward up to and including out_conv (mixlogcdf_nn.py:64-69): [B, (2+3K)*c, H, W]."""
    h = F.conv2d(x_id, wn(sd, pre + "in_conv.conv."), sd[pre + "in_conv.conv.bias"], padding=1)
    i = 0
    while (pre + "mid_convs.%d.norm_1.weight" % i) in sd:
        h = conv_attn_block(sd, pre + "mid_convs.%d." % i, h)
        i += 1
    return F.conv2d(h, wn(sd, pre + "out_conv.conv."), sd[pre + "out_conv.conv.bias"], padding=1)


def mixlogcdf_split_params(raw: Tensor, rescale_w: Tensor, k: int = 32):
    """Post-processing of the raw conditioner output (mixlogcdf_nn.py:72-78).

    raw viewed [B, 2+3K, c, H, W]; rows (0, 1, 2..2+K, 2+K..2+2K, 2+2K..2+3K) are
    (a_raw, b, pi, mu, s); a = rescale_w * tanh(a_raw); s = max(s, -7).
    """
    b, tot, h, w = raw.shape
    c = tot // (2 + 3 * k)
    r = raw.view(b, 2 + 3 * k, c, h, w)
    a = rescale_w * torch.tanh(r[:, 0])
    t = r[:, 1]
    pi, mu, s = r[:, 2:2 + k], r[:, 2 + k:2 + 2 * k], r[:, 2 + 2 * k:]
    return a, t, pi, mu, s.clamp(min=S_FLOOR)


def mixlogcdf_conditioner(sd: State, pre: str, x_id: Tensor, k: int = 32):
    """NN.forward (mixlogcdf_nn.py:64-78): returns (a, b, pi, mu, s)."""
    raw = mixlogcdf_conditioner_raw(sd, pre, x_id)
    return mixlogcdf_split_params(raw, wn(sd, pre + "rescale."), k)


# --------------------------------------------------------------------------
# logistic-mixture maths (flow_modules/log_dist.py)
# --------------------------------------------------------------------------
def _safe_log(x: Tensor) -> Tensor:
    return torch.log(x.clamp(min=LOG_FLOOR))                      # log_dist.py:5-6


def mix_log_cdf(x: Tensor, pi: Tensor, mu: Tensor, s: Tensor) -> Tensor:
    """log sum_k softmax(pi)_k sigmoid((x-mu_k) e^{-s_k})  (log_dist.py:17-22,34-40)."""
    z = (x.unsqueeze(1) - mu) * torch.exp(-s)
    return torch.logsumexp(F.log_softmax(pi, dim=1) + F.logsigmoid(z), dim=1)


def mix_log_pdf(x: Tensor, pi: Tensor, mu: Tensor, s: Tensor) -> Tensor:
    """log sum_k softmax(pi)_k logistic_pdf(x; mu_k, s_k)  (log_dist.py:9-14,25-31)."""
    z = (x.unsqueeze(1) - mu) * torch.exp(-s)
    return torch.logsumexp(F.log_softmax(pi, dim=1) + (z - s - 2 * F.softplus(z)), dim=1)


def mix_inv_cdf(y: Tensor, pi: Tensor, mu: Tensor, s: Tensor,
                eps: float = BISECT_EPS, max_iters: int = BISECT_MAX_ITERS,
                return_iters: bool = False):
    """Bisection inverse of the mixture CDF (log_dist.py:43-72).

    Start x=0, bracket [min_k(mu_k - 20 S), max_k(mu_k + 20 S)], S = sum_k e^{s_k};
    each step moves x to the midpoint with the bound on the side the CDF says,
    stops on a GLOBAL max|dx| <= eps or max_iters.
    """
    if y.min() <= 0 or y.max() >= 1:
        raise RuntimeError('Inverse logisitic CDF got y outside (0, 1)')   # log_dist.py:46-47
    x = torch.zeros_like(y)
    spread = torch.exp(s).sum(dim=1, keepdim=True)
    lb = (mu - 20 * spread).min(dim=1)[0]
    ub = (mu + 20 * spread).max(dim=1)[0]
    step = float('inf')
    it = 0
    while step > eps and it < max_iters:
        above = (torch.exp(mix_log_cdf(x, pi, mu, s)) > y).to(y.dtype)
        below = 1 - above
        nx = above * (x + lb) / 2. + below * (x + ub) / 2.
        lb = above * lb + below * x
        ub = above * x + below * ub
        step = (nx - x).abs().max()
        x = nx
        it += 1
    return (x, it) if return_iters else x


def mixlogcdf_elementwise(x: Tensor, a: Tensor, b: Tensor, pi: Tensor, mu: Tensor, s: Tensor,
                          ldj: Tensor, reverse: bool = False):
    """Coupling arithmetic alone, given conditioner outputs (mixlogcdf_coupling.py:41-57).

    FIRST half of x is transformed, second half passes through.  Returns cat(out, x_id).
    """
    c = x.shape[1] // 2
    xc, xid = x[:, :c], x[:, c:]
    if reverse:
        t = xc * torch.exp(-a) - b
        u = torch.sigmoid(t)                                        # log_dist.py:77-79
        scale_ldj = F.softplus(t) + F.softplus(-t)
        u = u.clamp(U_CLAMP, 1.0 - U_CLAMP)
        out = mix_inv_cdf(u, pi, mu, s)
        pdf_ldj = mix_log_pdf(out, pi, mu, s)
        ldj = ldj - (a + scale_ldj + pdf_ldj).flatten(1).sum(-1)
    else:
        u = mix_log_cdf(xc, pi, mu, s).exp()
        v = -_safe_log(u.reciprocal() - 1.0)                        # log_dist.py:81
        scale_ldj = -_safe_log(u) - _safe_log(1.0 - u)              # log_dist.py:82
        out = (v + b) * torch.exp(a)
        pdf_ldj = mix_log_pdf(xc, pi, mu, s)
        ldj = ldj + (pdf_ldj + scale_ldj + a).flatten(1).sum(-1)
    return torch.cat((out, xid), dim=1), ldj


def mixlogcdf_coupling(sd: State, pre: str, x: Tensor, ldj: Tensor, reverse: bool = False, k: int = 32):
    """MixLogCDFCoupling.forward (mixlogcdf_coupling.py:37-57); ``pre`` ends in 'coupling.'."""
    c = x.shape[1] // 2
    a, b, pi, mu, s = mixlogcdf_conditioner(sd, pre + "nn.", x[:, c:], k)
    return mixlogcdf_elementwise(x, a, b, pi, mu, s, ldj, reverse)


def tuple_flip(x: Tensor) -> Tensor:
    """Swap channel halves; its own inverse (common_modules.py:214-220)."""
    c = x.shape[1] // 2
    return torch.cat((x[:, c:], x[:, :c]), dim=1)


# --------------------------------------------------------------------------
# Transformer_attn, the fork's invertible patch attention (flow_modules/transformer.py:31-326;
# SURVEY.md section 8f-2).  Not on the north-star path; pinned for the optional plug-in.
# --------------------------------------------------------------------------
def _patches(x: Tensor, p: int) -> Tensor:
    """'b c (h p1) (w p2) -> b (h w) (c p1 p2)' (transformer.py:132)."""
    b, c, hh, ww = x.shape
    return x.reshape(b, c, hh // p, p, ww // p, p).permute(0, 2, 4, 1, 3, 5).reshape(b, (hh // p) * (ww // p), c * p * p)


def _unpatches(t: Tensor, p: int, shape) -> Tensor:
    """Inverse of `_patches` (reverse_rearrange, transformer.py:14-29)."""
    b, c, hh, ww = shape
    return t.reshape(b, hh // p, ww // p, c, p, p).permute(0, 3, 1, 4, 2, 5).reshape(b, c, hh, ww)


def _checker(rows: int, cols: int, like: Tensor) -> Tensor:
    """checkerboard(shape) = 1 - (i + j) % 2 (transformer.py:10-11)."""
    i = torch.arange(rows).view(-1, 1)
    j = torch.arange(cols).view(1, -1)
    return (1 - (i + j) % 2).to(like.dtype)


def transformer_attn(sd: State, pre: str, z: Tensor, ldj: Tensor, reverse: bool = False, permute: bool = False):
    """Masked patch attention with 2x2 block-invertible mixing (transformer.py:124-326).

    Patches: a 2x2 grid of (W/2)-sized patches, flattened to [B, 4, L].  Entries with (patch + index) even (odd when
    `permute`) condition the attention and pass through; the others are mixed between patches of equal parity by the
    2x2 matrices M1 (patches 0, 2) and M2 (patches 1, 3): attn = (sigmoid(sum_i Q_i K_i^T / scale + offset2) +
    offset3), plus `offset` on the diagonal.  logdet += (log|det M1| + log|det M2|) * p * (p // 2) * C."""
    b, c, hh, ww = z.shape
    p = ww // 2
    full = _patches(z, p)                                      # [B, 4, L]
    n, length = full.shape[1], full.shape[2]
    mask = _checker(n, length, z)
    if permute:
        mask = 1 - mask
    z_m = _unpatches(full * mask, p, z.shape)
    score = 0
    for i in (1, 2, 3):
        q = _patches(F.conv2d(z_m, sd[pre + "convq%d" % i]), p)
        k = _patches(F.conv2d(z_m, sd[pre + "convk%d" % i]), p)
        score = score + torch.matmul(q, k.permute(0, 2, 1)) / sd[pre + "scale"]
    attn = torch.sigmoid(score + sd[pre + "offset2"]) + sd[pre + "offset3"]          # [B, 4, 4]; only equal-parity entries used
    off = sd[pre + "offset"].reshape(())
    eye = torch.eye(2, dtype=z.dtype)
    m1 = attn[:, 0::2, 0::2] + eye * off                       # patches (0, 2)
    m2 = attn[:, 1::2, 1::2] + eye * off                       # patches (1, 3)
    scale_ld = p * (p // 2) * c
    ld = (torch.slogdet(m1)[1] + torch.slogdet(m2)[1]) * scale_ld
    free = full * (1 - mask)
    if not reverse:
        mixed = torch.empty_like(full)
        mixed[:, 0::2] = torch.matmul(m1, free[:, 0::2])
        mixed[:, 1::2] = torch.matmul(m2, free[:, 1::2])
        ldj = ldj + ld
    else:
        mixed = torch.empty_like(full)
        mixed[:, 0::2] = torch.matmul(torch.inverse(m1), free[:, 0::2])
        mixed[:, 1::2] = torch.matmul(torch.inverse(m2), free[:, 1::2])
        ldj = ldj - ld
    out = mixed * (1 - mask) + full * mask
    return _unpatches(out, p, z.shape), ldj


# --------------------------------------------------------------------------
# FlowStep / FlowNet / MarScfFlow (marscf_main.py:35-206).  `attn=True` adds the fork's two
# Transformer_attn layers after the 1x1 conv (marscf_main.py:69-70, 89-90).
# --------------------------------------------------------------------------
def flow_step(sd: State, pre: str, x: Tensor, ldj: Tensor, coupling: str, reverse: bool = False, attn: bool = False):
    """actnorm -> invconv -> coupling (-> flip if mixlogcdf); exact mirror in reverse
    (marscf_main.py:62-106)."""
    an = (sd[pre + "actnormlayer.bias"], sd[pre + "actnormlayer.logs"])
    ic = tuple(sd[pre + "invert_1x1_layer." + n] for n in ("p", "l", "u", "sign_s", "log_s"))
    if not reverse:
        x, ldj = actnorm(x, *an, ldj, False)
        x, ldj = invconv(x, *ic, ldj, False)
        if attn:
            x, ldj = transformer_attn(sd, pre + "attn1.", x, ldj, False, False)
            x, ldj = transformer_attn(sd, pre + "attn2.", x, ldj, False, True)
        if coupling == "mixlogcdf":
            x, ldj = mixlogcdf_coupling(sd, pre + "coupling.", x, ldj, False)
            x = tuple_flip(x)
        else:
            x, ldj = affine_coupling(sd, pre + "coupling.", x, ldj, False)
    else:
        if coupling == "mixlogcdf":
            x = tuple_flip(x)
            x, ldj = mixlogcdf_coupling(sd, pre + "coupling.", x, ldj, True)
        else:
            x, ldj = affine_coupling(sd, pre + "coupling.", x, ldj, True)
        if attn:
            x, ldj = transformer_attn(sd, pre + "attn2.", x, ldj, True, True)
            x, ldj = transformer_attn(sd, pre + "attn1.", x, ldj, True, False)
        x, ldj = invconv(x, *ic, ldj, True)
        x, ldj = actnorm(x, *an, ldj, True)
    return x, ldj


def layer_plan(L: int, K: int) -> List[str]:
    """Layer order of FlowNet.__init__ (marscf_main.py:127-145): per level squeeze, K steps,
    split between levels."""
    plan: List[str] = []
    for lvl in range(L):
        plan.append("squeeze")
        plan.extend(["step"] * K)
        if lvl < L - 1:
            plan.append("split")
    return plan


def gaussian_logp(z: Tensor) -> Tensor:
    """Standard-normal log-likelihood per sample (GaussianDiag.logp with mean=0, logs=0;
    common_modules.py:226-233).  Stand-in for the mAR prior, which is outside the hot path."""
    ll = -0.5 * (z ** 2 + math.log(2 * math.pi))
    return seq_sum(ll, [1, 2, 3])


def flownet_encode(sd: State, x: Tensor, ldj: Tensor, L: int, K: int, coupling: str,
                   pre: str = "flow.layers.", attn: bool = False):
    """FlowNet.encode (marscf_main.py:156-165) without the prior term: returns
    (z_final, [z2 of every split, in order], flow logdet)."""
    outs: List[Tensor] = []
    for i, kind in enumerate(layer_plan(L, K)):
        if kind == "squeeze":
            x = squeeze2d(x)
        elif kind == "step":
            x, ldj = flow_step(sd, "%s%d." % (pre, i), x, ldj, coupling, False, attn=attn)
        else:
            c = x.shape[1] // 2
            outs.append(x[:, c:])
            x = x[:, :c]
    return x, outs, ldj


def flownet_decode(sd: State, z: Tensor, z2s: Sequence[Tensor], L: int, K: int, coupling: str,
                   pre: str = "flow.layers.", attn: bool = False):
    """FlowNet.decode (marscf_main.py:167-175) with the factored-out halves supplied by the
    caller instead of sampled from the prior.  logdet restarts from 0 at every layer in the
    reference (:174) and is discarded; here it is accumulated and returned for testing."""
    plan = layer_plan(L, K)
    ldj = torch.zeros(z.shape[0], dtype=z.dtype)
    z2s = list(z2s)
    for i in reversed(range(len(plan))):
        kind = plan[i]
        if kind == "split":
            z = torch.cat((z, z2s.pop()), dim=1)
        elif kind == "step":
            z, ldj = flow_step(sd, "%s%d." % (pre, i), z, ldj, coupling, True, attn=attn)
        else:
            z = unsqueeze2d(z)
    return z, ldj


def normal_flow(sd: State, x: Tensor, noise: Tensor, L: int, K: int, coupling: str, attn: bool = False):
    """MarScfFlow.normal_flow (marscf_main.py:192-206) with the dequantisation noise supplied
    (``noise`` ~ U[0,1), added as noise/256) and a standard-normal prior on every latent.

    Returns (z_final, z2 list, flow logdet [B], bits/dim [B]).
    """
    d = x.shape[1] * x.shape[2] * x.shape[3]
    z = x + noise * (1.0 / 256.0)
    ldj = torch.zeros(x.shape[0], dtype=x.dtype) + float(-math.log(256.0) * d)
    z, outs, ldj = flownet_encode(sd, z, ldj, L, K, coupling, attn=attn)
    objective = ldj + gaussian_logp(z)
    for o in outs:
        objective = objective + gaussian_logp(o)
    nll = (-objective) / float(math.log(2.0) * d)
    return z, outs, ldj, nll
