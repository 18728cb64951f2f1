"""TEST INFRASTRUCTURE ONLY.  `autograd.numpy.linalg` stand-in (see ../_box.py)."""
import numpy as _np
import torch as _torch
from numpy.linalg import cholesky, norm, svd  # noqa: F401

from .._box import Box, _t, is_box


def _dual(np_fn, torch_fn):
    def fn(*args, **kwargs):
        if is_box(args):
            return torch_fn(*args, **kwargs)
        return np_fn(*args, **kwargs)
    return fn


det = _dual(_np.linalg.det, lambda a: Box(_torch.linalg.det(_t(a))))
inv = _dual(_np.linalg.inv, lambda a: Box(_torch.linalg.inv(_t(a))))
solve = _dual(_np.linalg.solve, lambda a, b: Box(_torch.linalg.solve(_t(a), _t(b))))
eigvalsh = _dual(_np.linalg.eigvalsh, lambda a: Box(_torch.linalg.eigvalsh(_t(a))))


def _eigh_box(a):
    w, v = _torch.linalg.eigh(_t(a))
    return Box(w), Box(v)


eigh = _dual(_np.linalg.eigh, _eigh_box)


def _slogdet_box(a):
    s, l = _torch.linalg.slogdet(_t(a))
    return Box(s), Box(l)


slogdet = _dual(_np.linalg.slogdet, _slogdet_box)
