"""TEST INFRASTRUCTURE ONLY.  Stand-in for `autograd.numpy`: plain numpy when no
argument is a Box, torch float64 (differentiable) when one is.  See ../_box.py."""
import builtins as _b

import numpy as _np
import torch as _torch
from numpy import (  # noqa: F401  (non-differentiable helpers pass straight through)
    pi, inf, nan, newaxis, e, float64, float32, int64, bool_, ndarray, finfo,
    arange, linspace, eye, identity, tril_indices, argsort, argmin, argmax, argwhere,
    isnan, isinf, isfinite, logical_not, logical_and, logical_or, any, all, delete,
    nanmean, apply_along_axis, where, ceil, floor, full, empty, cov, round,
    conjugate, real, imag, cumsum, unique, sort, copy, allclose, negative, divide,
    subtract, add, log10, log2, power, diff, mod, clip, isscalar)
from numpy import fft  # noqa: F401

from .._box import Box, _t, is_box, unbox
from . import linalg, random  # noqa: F401


def _dual(np_fn, torch_fn):
    def fn(*args, **kwargs):
        if is_box(args) or is_box(tuple(kwargs.values())):
            return torch_fn(*args, **kwargs)
        return np_fn(*args, **kwargs)
    fn.__name__ = np_fn.__name__
    return fn


def _un(tf):
    return lambda x: Box(tf(_t(x)))


exp = _dual(_np.exp, _un(_torch.exp))
log = _dual(_np.log, _un(_torch.log))
log1p = _dual(_np.log1p, _un(_torch.log1p))
expm1 = _dual(_np.expm1, _un(_torch.expm1))
sqrt = _dual(_np.sqrt, _un(_torch.sqrt))
tanh = _dual(_np.tanh, _un(_torch.tanh))
abs = _dual(_np.abs, _un(_torch.abs))
square = _dual(_np.square, _un(_torch.square))
sign = _dual(_np.sign, _un(_torch.sign))
trace = _dual(_np.trace, _un(_torch.trace))


def _red(tf):
    def fn(x, axis=None, keepdims=False):
        x = _t(x)
        if axis is None:
            return Box(tf(x))
        r = tf(x, dim=axis, keepdim=keepdims)
        return Box(r if isinstance(r, _torch.Tensor) else r[0])
    return fn


sum = _dual(_np.sum, _red(_torch.sum))
mean = _dual(_np.mean, _red(_torch.mean))
max = _dual(_np.max, _red(_torch.amax))
min = _dual(_np.min, _red(_torch.amin))
var = _dual(_np.var, lambda x, axis=None: Box(
    _torch.var(_t(x), unbiased=False) if axis is None
    else _torch.var(_t(x), dim=axis, unbiased=False)))
logaddexp = _dual(_np.logaddexp, lambda a, b: Box(_torch.logaddexp(_t(a), _t(b))))
maximum = _dual(_np.maximum, lambda a, b: Box(_torch.maximum(_t(a), _t(b))))
multiply = _dual(_np.multiply, lambda a, b: Box(_t(a) * _t(b)))
inner = _dual(_np.inner, lambda a, b: Box(_torch.inner(_t(a), _t(b))))
outer = _dual(_np.outer, lambda a, b: Box(_torch.outer(_t(a), _t(b))))
dot = _dual(_np.dot, lambda a, b: Box(_torch.matmul(_t(a), _t(b))))
matmul = _dual(_np.matmul, lambda a, b: Box(_torch.matmul(_t(a), _t(b))))
diag = _dual(_np.diag, lambda a: Box(_torch.diag(_t(a))))
reshape = _dual(_np.reshape, lambda a, shape: Box(_t(a).reshape(shape)))
transpose = _dual(_np.transpose, lambda a: Box(_t(a).T))
expand_dims = _dual(_np.expand_dims, lambda a, axis: Box(_t(a).unsqueeze(axis)))
squeeze = _dual(_np.squeeze, lambda a, axis=None: Box(a.t.squeeze() if axis is None
                                                       else a.t.squeeze(axis)))
ones_like = _dual(_np.ones_like, lambda a: _np.ones(a.shape))
zeros_like = _dual(_np.zeros_like, lambda a: _np.zeros(a.shape))
concatenate = _dual(_np.concatenate, lambda seq, axis=0: Box(
    _torch.cat([_t(s) for s in seq], dim=axis)))
column_stack = _dual(_np.column_stack, lambda seq: Box(
    _torch.column_stack([_t(s) for s in seq])))
stack = _dual(_np.stack, lambda seq, axis=0: Box(
    _torch.stack([_t(s) for s in seq], dim=axis)))
atleast_2d = _dual(_np.atleast_2d, lambda a: Box(_torch.atleast_2d(_t(a))))
atleast_1d = _dual(_np.atleast_1d, lambda a: Box(_torch.atleast_1d(_t(a))))


def shape(a):
    return tuple(a.shape)


class _Acc(_np.ndarray):
    """ndarray whose in-place add accepts a Box: `acc = np.zeros(n); acc += traced` (the accumulator pattern of
    NeuralNet.forward / NVPFlow.f, approximations.py:419-428, :520-529) rebinds `acc` to the traced sum, as real
    autograd's ArrayBox arithmetic does.  Everything else behaves as a plain ndarray."""

    def __iadd__(self, other):
        if isinstance(other, Box):
            return other + _np.asarray(self)
        return _np.ndarray.__iadd__(self, other)

    def __isub__(self, other):
        if isinstance(other, Box):
            return (-other) + _np.asarray(self)
        return _np.ndarray.__isub__(self, other)


def zeros(*a, **k):
    return _np.zeros(*a, **k).view(_Acc)


def ones(*a, **k):
    return _np.ones(*a, **k)


def asarray(a, dtype=None):
    if isinstance(a, Box):
        return a
    return _np.asarray(a, dtype=dtype)


def array(a, dtype=None, **kw):
    if isinstance(a, Box):
        return a
    if isinstance(a, (list, tuple)) and is_box(a):
        return Box(_torch.stack([_t(v).to(_torch.float64) for v in a]))
    return _np.array(a, dtype=dtype, **kw)
