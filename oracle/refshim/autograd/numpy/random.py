"""TEST INFRASTRUCTURE ONLY.  `autograd.numpy.random` stand-in: the real numpy
legacy generators (so the reference draws exactly what it would draw), plus a
recording hook so golden vectors can store the base draws for injection."""
import numpy as _np
from numpy.random import (  # noqa: F401
    randint, seed, randn, rand, choice, multivariate_normal, normal, uniform)

#: when set to a list, every RandomState draw is appended as (method, args, ndarray)
RECORD = None
#: when set to a callable (method, args, kwargs) -> ndarray | None, overrides draws
INJECT = None


class RandomState(object):
    def __init__(self, seed=None):
        self._rs = _np.random.RandomState(seed)
        self.seed_used = seed

    def _draw(self, name, *args, **kwargs):
        out = None
        if INJECT is not None:
            out = INJECT(name, args, kwargs)
        if out is None:
            out = getattr(self._rs, name)(*args, **kwargs)
        if RECORD is not None:
            RECORD.append((name, args, _np.array(out, copy=True)))
        return out

    def randn(self, *a, **k):
        return self._draw('randn', *a, **k)

    def standard_t(self, *a, **k):
        return self._draw('standard_t', *a, **k)

    def chisquare(self, *a, **k):
        return self._draw('chisquare', *a, **k)

    def rand(self, *a, **k):
        return self._draw('rand', *a, **k)

    def __getattr__(self, name):
        return getattr(self._rs, name)
