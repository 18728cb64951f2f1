"""TEST INFRASTRUCTURE ONLY.  `autograd.scipy.special` stand-in."""
import scipy.special as _sp
import torch as _torch

from .._box import Box, _t, is_box


def gammaln(x):
    if is_box((x,)):
        return Box(_torch.lgamma(_t(x)))
    return _sp.gammaln(x)


def logsumexp(x, axis=None):
    if is_box((x,)):
        return Box(_torch.logsumexp(_t(x), dim=tuple(range(_t(x).dim())) if axis is None else axis))
    return _sp.logsumexp(x, axis=axis)
