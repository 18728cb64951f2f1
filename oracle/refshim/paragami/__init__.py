"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-in for `paragami` (reference requirements.txt:5, paragami~=0.42; not in this
image).  Restates the published behaviour of the four pattern classes the
reference uses (approximations.py:9-11):

* PatternDict(free_default=True): flat vector = concatenation of the members'
  flat vectors in insertion order.
* NumericVectorPattern(length) / NumericArrayPattern(shape): unconstrained,
  flat = C-order ravel.
* PSDSymmetricMatrixPattern(size), diag_lb=0: free vector = lower triangle of the
  Cholesky factor of the matrix, row-major (np.tril_indices order), with the log
  taken on the diagonal.
* FlattenFunctionInput(f, patterns, free, argnums=0): f(fold(flat), *rest).

NOTE (parity): no reference test pins this layout; SURVEY.md 8(c) records the
evidence (FASO slices [:dim]/[-dim:], MultivariateT.init_param probe).
"""
import numpy as _np
import torch as _torch

from autograd._box import Box, _t


class NumericArrayPattern(object):
    def __init__(self, shape, **kw):
        self._shape = tuple(shape)

    def flat_length(self, free=None):
        return int(_np.prod(self._shape))

    def flatten(self, val, free=None):
        return _np.asarray(val, dtype=float).reshape(-1)

    def fold(self, flat, free=None):
        return flat.reshape(self._shape)


class NumericVectorPattern(NumericArrayPattern):
    def __init__(self, length, **kw):
        super().__init__((length,))


class PSDSymmetricMatrixPattern(object):
    def __init__(self, size, diag_lb=0.0, **kw):
        self._n = size
        self._lb = diag_lb

    def flat_length(self, free=None):
        return self._n * (self._n + 1) // 2

    def flatten(self, val, free=None):
        L = _np.linalg.cholesky(_np.asarray(val, dtype=float) - self._lb * _np.eye(self._n))
        L = L.copy()
        L[_np.diag_indices(self._n)] = _np.log(_np.diag(L))
        return L[_np.tril_indices(self._n)]

    def fold(self, flat, free=None):
        n = self._n
        r, c = _np.tril_indices(n)
        if isinstance(flat, Box):
            F = _torch.zeros((n, n), dtype=_torch.float64)
            F = F.index_put((_torch.from_numpy(r), _torch.from_numpy(c)), flat.t)
            L = _torch.tril(F, -1) + _torch.diag(_torch.exp(_torch.diagonal(F)))
            return Box(L @ L.T + self._lb * _torch.eye(n, dtype=_torch.float64))
        F = _np.zeros((n, n))
        F[r, c] = flat
        L = _np.tril(F, -1) + _np.diag(_np.exp(_np.diag(F)))
        return L @ L.T + self._lb * _np.eye(n)


class PatternDict(object):
    def __init__(self, free_default=None):
        self._free = free_default
        self._keys = []
        self._pats = {}

    def __setitem__(self, key, pat):
        if key not in self._pats:
            self._keys.append(key)
        self._pats[key] = pat

    def __getitem__(self, key):
        return self._pats[key]

    def flat_length(self, free=None):
        return sum(self._pats[k].flat_length(free) for k in self._keys)

    def flatten(self, d, free=None):
        return _np.concatenate([self._pats[k].flatten(d[k], free) for k in self._keys])

    def fold(self, flat, free=None):
        out, off = {}, 0
        for k in self._keys:
            n = self._pats[k].flat_length(free)
            out[k] = self._pats[k].fold(flat[off:off + n], free)
            off += n
        return out


class FlattenFunctionInput(object):
    def __init__(self, original_fun, patterns, free, argnums=0):
        assert argnums == 0
        self._f, self._p, self._free = original_fun, patterns, free

    def __call__(self, flat, *args, **kwargs):
        return self._f(self._p.fold(flat, self._free), *args, **kwargs)
