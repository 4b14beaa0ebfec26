"""Flow primitives with the reference's class names, constructor arguments, state-dict keys and
`forward(x, logdet, reverse)` contract (flow_modules/common_modules.py:12-240), computed by the
sm_100a kernels in libflowk.so through `flowk.ops`.

Differences a caller can observe:
  * float `logdet` arguments (the reference's default `0.`) come back as a [B] tensor, never 0-dim;
  * InvertibleConv1x1 inverts L and U in float64 on the device (the reference round-trips through
    the CPU, common_modules.py:108-110) and caches W / W^-1 per parameter version in no-grad mode;
  * Actnormlayer's "is it initialised yet" test syncs with the host only until it has been seen true.
"""
import numpy as np
import scipy.linalg
import torch
import torch.nn as nn

from .. import _lib, ops
from .misc import cpd_sum


def _batch_ldj(logdet, x):
    """Normalise the reference's polymorphic logdet argument (None | float | tensor) to ([B] tensor, had_ldj)."""
    if logdet is None:
        return x.new_zeros(x.shape[0]), False
    if not torch.is_tensor(logdet):
        return x.new_full((x.shape[0],), float(logdet)), True
    if logdet.dim() == 0:
        return logdet.to(x.dtype).expand(x.shape[0]).contiguous(), True
    return logdet, True


def squeeze2d(input, factor=2):
    if factor == 1:
        return input
    h, w = input.shape[2], input.shape[3]
    assert h % factor == 0 and w % factor == 0, "{}".format((h, w))
    return ops.squeeze2d(input, factor)


def unsqueeze2d(input, factor=2):
    assert factor >= 1 and isinstance(factor, int)
    if factor == 1:
        return input
    assert input.shape[1] % (factor * factor) == 0, "{}".format(input.shape[1])
    return ops.unsqueeze2d(input, factor)


class SqueezeLayer(nn.Module):
    def __init__(self, factor):
        super().__init__()
        self.factor = factor

    def forward(self, input, logdet=0., reverse=False):
        if not reverse:
            return squeeze2d(input, self.factor), logdet
        return unsqueeze2d(input, self.factor), logdet


class Actnormlayer(nn.Module):
    """y = (x + bias) * exp(logs), data-dependent init on the first training batch."""

    def __init__(self, num_features, scale=1.):
        super().__init__()
        self.register_buffer('is_initialized', torch.zeros(1))
        self.bias = nn.Parameter(torch.zeros(1, num_features, 1, 1))
        self.logs = nn.Parameter(torch.zeros(1, num_features, 1, 1))
        self.num_features = num_features
        self.scale = float(scale)
        self.eps = 1e-6
        self._seen_initialized = False

    def _load_from_state_dict(self, *args, **kwargs):
        self._seen_initialized = False
        return super()._load_from_state_dict(*args, **kwargs)

    def initialize_parameters(self, x):
        if not self.training:
            return
        with torch.no_grad():
            bias, logs = ops.actnorm_init(x.detach(), self.scale, self.eps)
            self.bias.data.copy_(bias.view_as(self.bias))
            self.logs.data.copy_(logs.view_as(self.logs))
            self.is_initialized += 1.
        _lib.bump_generation()        # `.data` writes skip the version counter

    def maybe_initialize(self, x):
        """The reference tests `if not self.is_initialized` every call (a host sync); here the
        answer is cached once it has been observed true.  Nothing happens outside train()."""
        if self._seen_initialized or not self.training:
            return
        if not bool(self.is_initialized):
            self.initialize_parameters(x)
        self._seen_initialized = True

    def ldj_term(self, x):
        return self.logs.sum() * (x.size(2) * x.size(3))

    def forward(self, x, ldj=None, reverse=False):
        self.maybe_initialize(x)
        ldj_t, had = _batch_ldj(ldj, x)
        logs = self.logs.view(-1)
        bias = self.bias.view(-1)
        zero = torch.zeros_like(bias)
        d = self.ldj_term(x).reshape(1)
        if reverse:
            y, out = ops.channel_scale(x, zero, torch.exp(-logs), -bias, ldj_t, -d)
        else:
            y, out = ops.channel_scale(x, bias, torch.exp(logs), zero, ldj_t, d)
        return y, (out if had else None)


class InvertibleConv1x1(nn.Module):
    """LU-parametrised invertible 1x1 convolution (common_modules.py:57-127)."""

    def __init__(self, num_channels, LU_decomposed=True):
        super().__init__()
        w_shape = [num_channels, num_channels]
        w_init = np.linalg.qr(np.random.randn(*w_shape))[0].astype(np.float32)
        if not LU_decomposed:
            self.weight = nn.Parameter(torch.from_numpy(w_init))
        else:
            perm, lower, upper = scipy.linalg.lu(w_init)
            diag = np.diag(upper)
            self.register_buffer('p', torch.from_numpy(perm.astype(np.float32)))
            self.register_buffer('sign_s', torch.from_numpy(np.sign(diag).astype(np.float32)))
            self.l = nn.Parameter(torch.from_numpy(lower.astype(np.float32)))
            self.log_s = nn.Parameter(torch.from_numpy(np.log(np.abs(diag)).astype(np.float32)))
            self.u = nn.Parameter(torch.from_numpy(np.triu(upper, k=1).astype(np.float32)))
        self.w_shape = w_shape
        self.LU = LU_decomposed
        self._cache = {}

    # -- weight assembly -----------------------------------------------------------------------
    def _params(self):
        return (self.l, self.u, self.log_s, self.p, self.sign_s) if self.LU else (self.weight,)

    def _build(self, reverse):
        c = self.w_shape[0]
        if not self.LU:
            logabsdet = torch.slogdet(self.weight)[1]
            w = self.weight if not reverse else torch.inverse(self.weight.double()).float()
            return w, logabsdet
        dev, dt = self.l.device, self.l.dtype
        lower_mask = torch.tril(torch.ones(c, c, device=dev, dtype=dt), -1)
        eye = torch.eye(c, device=dev, dtype=dt)
        lo = self.l * lower_mask + eye
        up = self.u * lower_mask.t() + torch.diag(self.sign_s * torch.exp(self.log_s))
        if not reverse:
            w = self.p @ (lo @ up)
        else:
            lo_inv = torch.linalg.solve_triangular(lo.double(), eye.double(), upper=False).to(dt)
            up_inv = torch.linalg.solve_triangular(up.double(), eye.double(), upper=True).to(dt)
            w = up_inv @ (lo_inv @ self.p.t())          # P is a permutation: P^-1 = P^T
        return w, cpd_sum(self.log_s)

    def weight_and_logabsdet(self, reverse):
        """(W or W^-1 as [C,C], sum log|s|).  Cached per parameter version when autograd is off."""
        params = self._params()
        track = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if track:
            return self._build(reverse)
        key = (bool(reverse),) + _lib.param_key(params)
        hit = self._cache.get(bool(reverse))
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, self._build(reverse))
            self._cache[bool(reverse)] = hit
        return hit[1]

    def get_weight(self, input, reverse):
        w, logabsdet = self.weight_and_logabsdet(reverse)
        pixels = input.size(-1)                      # last spatial dim, squared: common_modules.py:86,104
        c = self.w_shape[0]
        return w.view(c, c, 1, 1), logabsdet * pixels * pixels

    def forward(self, input, logdet=None, reverse=False):
        weight, dlogdet = self.get_weight(input, reverse)
        ldj_t, had = _batch_ldj(logdet, input)
        c = self.w_shape[0]
        d = dlogdet.reshape(1)
        z, out = ops.channel_mix(input, weight.view(c, c), None, ldj_t, -d if reverse else d, False, False)
        return z, (out if had else None)


class Split2dMsC(nn.Module):
    """Channel split feeding the multi-scale prior: views only (common_modules.py:189-208)."""

    def __init__(self, num_channels, level=0):
        super().__init__()
        self.level = level

    def split_feature(self, z):
        c = z.size(1) // 2
        return z[:, :c], z[:, c:]

    def forward(self, input, logdet=0., reverse=False, eps_std=None):
        if not reverse:
            return self.split_feature(input), logdet
        z1, z2 = input
        return torch.cat((z1, z2), dim=1), logdet


class TupleFlip(nn.Module):
    """Swap the channel halves (its own inverse).  Inside FlowStep the swap is fused into the
    MixLogCDF kernel's store; this module is the stand-alone equivalent."""

    def forward(self, z, logdet=0., reverse=False):
        a, b = z.chunk(2, dim=1)
        return torch.cat([b, a], dim=1), logdet


class GaussianDiag:
    Log2PI = float(np.log(2 * np.pi))

    @staticmethod
    def likelihood(mean, logs, x):
        return -0.5 * (logs * 2. + ((x - mean) ** 2) / torch.exp(logs * 2.) + GaussianDiag.Log2PI)

    @staticmethod
    def logp(mean, logs, x):
        return cpd_sum(GaussianDiag.likelihood(mean, logs, x), dim=[1, 2, 3])

    @staticmethod
    def sample(mean, logs, eps_std=None):
        eps_std = eps_std or 1
        eps = torch.normal(mean=torch.zeros_like(mean), std=torch.ones_like(logs) * eps_std)
        return mean + torch.exp(logs) * eps


FOLD_KERNEL = True      # False: assemble the fold with torch ops (tests A/B the two)


def fold_actnorm_invconv(actnorm, invconv, x_hw, reverse):
    """ActNorm followed by the 1x1 conv (or their inverses in reverse order) as ONE per-pixel affine
    map, so FlowStep needs a single pass over the activations (marscf_main.py:64-68 / :95-97):

        forward:  W (x + b) e^{logs}           = (W diag(e^{logs})) x + (W diag(e^{logs})) b
        reverse:  (W^-1 y) e^{-logs} - b       = (diag(e^{-logs}) W^-1) y - b

    Returns (matrix [C,C], bias [C], ldj_add [1]); ldj_add already carries the reverse sign and both
    log-det terms: sum(logs) H W (common_modules.py:167) and sum(log_s) W^2 (common_modules.py:104).
    """
    h, w = x_hw
    c = invconv.w_shape[0]
    params = (actnorm.bias, actnorm.logs) + tuple(invconv._params())
    if (FOLD_KERNEL and invconv.LU and c <= 158 and all(t.is_cuda and t.dtype == torch.float32 for t in params)
            and not (torch.is_grad_enabled() and any(t.requires_grad for t in params))):
        # inference / sampling: the whole assembly (masks, triangular inverses in fp64, products, fold) is one flowk launch
        _lib.check_device(invconv.l, "fold_actnorm_invconv")
        dev = invconv.l.device
        mat = torch.empty(c, c, device=dev, dtype=torch.float32)
        out = torch.empty(c + 1, device=dev, dtype=torch.float32)
        ptr = lambda t: t.detach().contiguous().data_ptr()       # noqa: E731 (parameters / buffers are contiguous: no copy)
        _lib.call("flowk_fold_actnorm_invconv", ptr(invconv.l), ptr(invconv.u), ptr(invconv.log_s), ptr(invconv.p),
                  ptr(invconv.sign_s), ptr(actnorm.logs), ptr(actnorm.bias), c, int(h), int(w), int(bool(reverse)),
                  mat.data_ptr(), out.data_ptr(), out[c:].data_ptr(), torch.cuda.current_stream().cuda_stream)
        return mat, out[:c], out[c:]
    mat, logabsdet = invconv.weight_and_logabsdet(reverse)
    logs = actnorm.logs.view(-1)
    bias = actnorm.bias.view(-1)
    d = logs.sum() * (h * w) + logabsdet * (w * w)
    if not reverse:
        m = mat * torch.exp(logs).unsqueeze(0)
        return m, m @ bias, d.reshape(1)
    m = torch.exp(-logs).unsqueeze(1) * mat
    return m, -bias, (-d).reshape(1)
