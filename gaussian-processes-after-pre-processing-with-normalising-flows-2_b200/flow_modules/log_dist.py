"""Logistic-mixture functions with the reference's names (flow_modules/log_dist.py:5-84), on the
flowk kernels.  Parameter tensors are [B, K, *x.shape[1:]] as in the reference."""
import torch
import torch.nn.functional as F

from .. import ops


def safe_log(x):
    return torch.log(x.clamp(min=1e-22))


def mixture_log_pdf(x, prior_logits, means, log_scales):
    return ops.mixture_log_pdf(x, prior_logits, means, log_scales)


def mixture_log_cdf(x, prior_logits, means, log_scales):
    return ops.mixture_log_cdf(x, prior_logits, means, log_scales)


def mixture_inv_cdf(y, prior_logits, means, log_scales, eps=1e-10, max_iters=100):
    """Bisection inverse; eps and max_iters are the reference's defaults and are compiled in."""
    if eps != 1e-10 or max_iters != 100:
        raise ValueError("flowk builds the reference's eps=1e-10, max_iters=100 only")
    if y.min() <= 0 or y.max() >= 1:
        raise RuntimeError('Inverse logisitic CDF got y outside (0, 1)')
    return ops.mixture_inv_cdf(y, prior_logits, means, log_scales)


def inverse(x, reverse=False):
    """Logit / sigmoid with their log-derivatives (log_dist.py:75-84)."""
    if reverse:
        return torch.sigmoid(x), F.softplus(x) + F.softplus(-x)
    return -safe_log(x.reciprocal() - 1.), -safe_log(x) - safe_log(1. - x)
