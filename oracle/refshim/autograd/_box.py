"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A minimal "numpy array backed by a torch float64 tensor" box.  It exists so the
UNMODIFIED reference sources under /root/reference/viabel (which are written
against `autograd.numpy` and differentiate with `autograd.value_and_grad`) can
be executed in this container, where the real `autograd` package is not
installed: torch's reverse-mode AD plays the role of autograd's tape.

Only the numpy surface the reference actually touches is covered
(grep over viabel/{approximations,objectives,_distributions,models}.py).
"""
import numpy as _np
import torch as _torch

_F64 = _torch.float64


def _t(x):
    """Coerce ndarray / scalar / Box to a torch float64 tensor."""
    if isinstance(x, Box):
        return x.t
    if isinstance(x, _torch.Tensor):
        return x
    a = _np.asarray(x)
    # np.array(copy=True) keeps 0-d arrays 0-d (ascontiguousarray would promote to 1-d)
    if a.dtype == _np.bool_:
        return _torch.from_numpy(_np.array(a, copy=True))
    if a.dtype.kind in 'iu':
        return _torch.from_numpy(_np.array(a, dtype=_np.int64, copy=True))
    return _torch.from_numpy(_np.array(a, dtype=_np.float64, copy=True))


def is_box(x):
    if isinstance(x, Box):
        return True
    if isinstance(x, (list, tuple)):
        return any(is_box(v) for v in x)
    return False


def unbox(x):
    """Box -> numpy (detached); everything else unchanged."""
    if isinstance(x, Box):
        return x.t.detach().numpy().copy()
    return x


class Box(object):
    __array_ufunc__ = None       # ndarray (op) Box -> Box.__r<op>__
    __array_priority__ = 10000

    def __init__(self, t):
        self.t = t

    # ---- array attributes -------------------------------------------------
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return _np.dtype('float64')

    @property
    def T(self):
        return Box(self.t.transpose(-1, -2) if self.t.dim() >= 2 else self.t)

    def __len__(self):
        return self.t.shape[0]

    def __float__(self):
        return float(self.t.detach())

    def __iter__(self):
        for i in range(self.t.shape[0]):
            yield Box(self.t[i])

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(_t(i) if isinstance(i, (_np.ndarray, Box)) else i for i in idx)
        elif isinstance(idx, (_np.ndarray, Box)):
            idx = _t(idx)
        return Box(self.t[idx])

    def copy(self):
        return Box(self.t.clone())

    def squeeze(self, axis=None):
        return Box(self.t.squeeze() if axis is None else self.t.squeeze(axis))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Box(self.t.reshape(shape))

    def sum(self, axis=None):
        return Box(self.t.sum() if axis is None else self.t.sum(dim=axis))

    # ---- arithmetic -------------------------------------------------------
    def __neg__(self):
        return Box(-self.t)

    def __add__(self, o):
        return Box(self.t + _t(o))
    __radd__ = __add__

    def __sub__(self, o):
        return Box(self.t - _t(o))

    def __rsub__(self, o):
        return Box(_t(o) - self.t)

    def __mul__(self, o):
        return Box(self.t * _t(o))
    __rmul__ = __mul__

    def __truediv__(self, o):
        return Box(self.t / _t(o))

    def __rtruediv__(self, o):
        return Box(_t(o) / self.t)

    def __pow__(self, o):
        return Box(self.t ** _t(o))

    def __rpow__(self, o):
        return Box(_t(o) ** self.t)

    def __matmul__(self, o):
        return Box(self.t @ _t(o))

    def __rmatmul__(self, o):
        return Box(_t(o) @ self.t)

    def __abs__(self):
        return Box(self.t.abs())

    # comparisons give plain numpy booleans (not differentiable)
    def __lt__(self, o):
        return (self.t.detach() < _t(o)).numpy()

    def __le__(self, o):
        return (self.t.detach() <= _t(o)).numpy()

    def __gt__(self, o):
        return (self.t.detach() > _t(o)).numpy()

    def __ge__(self, o):
        return (self.t.detach() >= _t(o)).numpy()

    def __eq__(self, o):
        return (self.t.detach() == _t(o)).numpy()

    __hash__ = None

    def __bool__(self):
        return bool(self.t.detach())

    def __repr__(self):
        return 'Box(%r)' % (self.t,)
