"""TEST INFRASTRUCTURE ONLY.  `autograd.scipy.stats.norm` stand-in."""
import math as _m

import scipy.stats as _st
import torch as _torch

from ..._box import Box, _t, is_box


def logpdf(x, loc=0.0, scale=1.0):
    if is_box((x, loc, scale)):
        x, loc, scale = _t(x), _t(loc), _t(scale)
        z = (x - loc) / scale
        return Box(-0.5 * z * z - _torch.log(scale) - 0.5 * _m.log(2 * _m.pi))
    return _st.norm.logpdf(x, loc, scale)


def pdf(x, loc=0.0, scale=1.0):
    if is_box((x, loc, scale)):
        return Box(_torch.exp(logpdf(x, loc, scale).t))
    return _st.norm.pdf(x, loc, scale)


def cdf(x, loc=0.0, scale=1.0):
    if is_box((x, loc, scale)):
        x, loc, scale = _t(x), _t(loc), _t(scale)
        return Box(0.5 * _torch.erfc(-(x - loc) / scale / _m.sqrt(2.0)))
    return _st.norm.cdf(x, loc, scale)


def logcdf(x, loc=0.0, scale=1.0):
    if is_box((x, loc, scale)):
        x, loc, scale = _t(x), _t(loc), _t(scale)
        return Box(_torch.special.log_ndtr((x - loc) / scale))
    return _st.norm.logcdf(x, loc, scale)
