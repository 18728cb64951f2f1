"""TEST INFRASTRUCTURE ONLY.  `autograd.scipy.stats.t` stand-in."""
import math as _m

import scipy.stats as _st
import torch as _torch

from ..._box import Box, _t, is_box


def logpdf(x, df, loc=0.0, scale=1.0):
    if is_box((x, df, loc, scale)):
        x, df, loc, scale = _t(x), _t(df).to(_torch.float64), _t(loc), _t(scale)
        z = (x - loc) / scale
        c = (_torch.lgamma(0.5 * (df + 1.0)) - _torch.lgamma(0.5 * df)
             - 0.5 * _torch.log(df * _m.pi))
        return Box(c - 0.5 * (df + 1.0) * _torch.log1p(z * z / df) - _torch.log(scale))
    return _st.t.logpdf(x, df, loc, scale)
