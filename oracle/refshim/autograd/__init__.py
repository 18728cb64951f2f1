"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-in for the `autograd` package (absent from this image; requirements.txt:3 of
the reference pins autograd~=1.3) so the unmodified reference sources can run
here: reverse-mode AD is delegated to torch (float64, CPU).  Used only by
oracle/make_golden.py and by tests that are skipped when /root/reference is absent."""
import numpy as _np
import torch as _torch

from ._box import Box, _t, unbox


def _lift(x):
    return Box(_torch.tensor(_np.asarray(x, dtype=_np.float64), requires_grad=True))


def value_and_grad(fun, argnum=0):
    def vg(*args, **kwargs):
        args = list(args)
        x = _lift(unbox(args[argnum]))
        args[argnum] = x
        out = fun(*args, **kwargs)
        if not isinstance(out, Box):            # constant function
            return float(out), _np.zeros(x.shape)
        (g,) = _torch.autograd.grad(out.t, x.t, allow_unused=True)
        g = _np.zeros(x.shape) if g is None else g.numpy().copy()
        return float(out.t.detach()), g
    return vg


def grad(fun, argnum=0):
    vg = value_and_grad(fun, argnum)
    return lambda *a, **k: vg(*a, **k)[1]


def elementwise_grad(fun, argnum=0):
    def eg(*args, **kwargs):
        args = list(args)
        x = _lift(unbox(args[argnum]))
        args[argnum] = x
        out = fun(*args, **kwargs)
        (g,) = _torch.autograd.grad(out.t.sum(), x.t)
        return g.numpy().copy()
    return eg


def vector_jacobian_product(fun, argnum=0):
    """autograd convention: the cotangent vector is the LAST positional argument."""
    def vjp(*args, **kwargs):
        args, vec = list(args[:-1]), args[-1]
        x = _lift(unbox(args[argnum]))
        args[argnum] = x
        out = fun(*args, **kwargs)
        (g,) = _torch.autograd.grad(out.t, x.t, grad_outputs=_t(unbox(vec)))
        return g.numpy().copy()
    return vjp


def hessian(fun, argnum=0):
    def h(*args, **kwargs):
        args = list(args)
        x0 = _torch.tensor(_np.asarray(unbox(args[argnum]), dtype=_np.float64))

        def f(xt):
            a = list(args)
            a[argnum] = Box(xt)
            return fun(*a, **kwargs).t.reshape(())
        return _torch.autograd.functional.hessian(f, x0).numpy().copy()
    return h


def make_hvp(fun, argnum=0):
    def at(*args, **kwargs):
        args = list(args)
        x0 = _torch.tensor(_np.asarray(unbox(args[argnum]), dtype=_np.float64))

        def f(xt):
            a = list(args)
            a[argnum] = Box(xt)
            return fun(*a, **kwargs).t.reshape(())

        def hvp(v):
            _, r = _torch.autograd.functional.hvp(
                f, x0, _torch.tensor(_np.asarray(unbox(v), dtype=_np.float64)))
            return r.numpy().copy()
        return (hvp,)
    return at
