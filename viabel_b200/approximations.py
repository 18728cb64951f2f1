"""Variational families -- drop-in mirror of viabel/approximations.py for the in-scope
families (MFGaussian :192-251, MFStudentT :254-312, MultivariateT :322-382).

Sampling and log densities run as CUDA kernels through the C ABI; the O(d) closed forms
(entropy, kl, moments) are evaluated on the host in float64.  Base draws come from a
counter-based Philox generator, not numpy's MT19937 (parity is by draw injection: pass
`base=` to `sample`, or read `last_base` after a call).
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _lib
from ._tensor import F64, device, is_host, like_input, to_dev

__all__ = ['ApproximationFamily', 'MFGaussian', 'MFStudentT']


class ApproximationFamily(ABC):
    """approximations.py:26-182"""

    def __init__(self, dim, var_param_dim, supports_entropy, supports_kl):
        self._dim = dim
        self._var_param_dim = var_param_dim
        self._supports_entropy = supports_entropy
        self._supports_kl = supports_kl

    def init_param(self):
        return np.zeros(self.var_param_dim)

    @abstractmethod
    def sample(self, var_param, n_samples, seed=None):
        """Generate samples from the variational distribution, shape (n_samples, dim)."""

    def entropy(self, var_param):
        if self.supports_entropy:
            return self._entropy(var_param)
        raise NotImplementedError()

    def _entropy(self, var_param):
        raise NotImplementedError()

    @property
    def supports_entropy(self):
        return self._supports_entropy

    def kl(self, var_param0, var_param1):
        if self.supports_kl:
            return self._kl(var_param0, var_param1)
        raise NotImplementedError()

    def _kl(self, var_param0, var_param1):
        raise NotImplementedError()

    @property
    def supports_kl(self):
        return self._supports_kl

    @abstractmethod
    def log_density(self, var_param, x):
        """Log density of the variational distribution at x."""

    @abstractmethod
    def mean_and_cov(self, var_param):
        """Mean and covariance of the variational distribution."""

    def pth_moment(self, var_param, p):
        if self.supports_pth_moment(p):
            return self._pth_moment(var_param, p)
        raise ValueError('p = {} is not a supported moment'.format(p))

    @abstractmethod
    def _pth_moment(self, var_param, p):
        """pth moment"""

    @abstractmethod
    def supports_pth_moment(self, p):
        """Whether the pth moment is available in closed form."""

    @property
    def dim(self):
        return self._dim

    @property
    def var_param_dim(self):
        return self._var_param_dim


def _host(var_param):
    if isinstance(var_param, torch.Tensor):
        return var_param.detach().cpu().numpy().astype(np.float64)
    return np.asarray(var_param, dtype=np.float64)


class _MeanField(ApproximationFamily):
    """Shared machinery of the two mean-field families; var_param = [mu, log_sigma]."""
    _family = None

    def __init__(self, dim, supports_kl, seed):
        self._seed = int(seed)
        self._offset = 0
        self.quantize_draws = 0          # 1: bf16-exact, 2: fp16-exact draws (exact tensor-core operands)
        self.last_base = None
        super().__init__(dim, 2 * dim, True, supports_kl)

    @property
    def df(self):
        return getattr(self, '_df', float('inf'))

    def init_param(self):
        # approximations.py:207-210, :265-268
        return np.concatenate([np.zeros(self.dim), 2.0 * np.ones(self.dim)])

    # -- base draws ---------------------------------------------------------------------------
    def _draw(self, n, seed, offset):
        raise NotImplementedError

    def base_draws(self, n_samples, seed=None):
        """Draw (n_samples, dim) base variates on the device.  With seed=None the family's own
        stream advances (like approx._rs); an explicit seed restarts a fresh stream at 0
        (like RandomState(seed), approximations.py:213)."""
        n = int(n_samples) * self.dim
        if seed is None:
            out = self._draw(n, self._seed, self._offset)
            self._offset += n + (n & 1)
        else:
            out = self._draw(n, int(seed), 0)
        return out.view(int(n_samples), self.dim)

    def sample(self, var_param, n_samples, seed=None, base=None):
        host = is_host(var_param)
        vp = to_dev(var_param)
        if vp.numel() != self.var_param_dim:
            raise ValueError('var_param has the wrong length')
        e = self.base_draws(n_samples, seed) if base is None else to_dev(base)
        self.last_base = e
        theta = torch.empty_like(e)
        _lib.check(_lib.lib.vb_mf_sample_f64(_lib.ptr(vp), _lib.ptr(e), _lib.ptr(theta), e.shape[0],
                                              self.dim, _lib.stream()))
        return like_input(theta, host)

    def log_density(self, var_param, x):
        host = is_host(x)
        vp = to_dev(var_param)
        xd = to_dev(x)
        if xd.dim() == 1:
            xd = xd[None, :]
        out = torch.empty(xd.shape[0], dtype=F64, device=device())
        _lib.check(_lib.lib.vb_mf_log_density_f64(_lib.ptr(vp), _lib.ptr(xd), xd.shape[0], self.dim,
                                                   self._family, float(self.df) if self._family else 0.0,
                                                   _lib.ptr(out), _lib.stream()))
        return like_input(out, host)


class MFGaussian(_MeanField):
    """A mean-field Gaussian approximation family (approximations.py:192-251)."""
    _family = _lib.FAMILY_MF_GAUSSIAN

    def __init__(self, dim, seed=1):
        super().__init__(dim, True, seed)

    def _draw(self, n, seed, offset):
        out = torch.empty(n, dtype=F64, device=device())
        _lib.check(_lib.lib.vb_philox_normal_f64(_lib.ptr(out), n, seed, offset,
                                                  int(self.quantize_draws), _lib.stream()))
        return out

    def _entropy(self, var_param):
        vp = _host(var_param)
        return 0.5 * self.dim * (1.0 + np.log(2 * np.pi)) + np.sum(vp[self.dim:])

    def _kl(self, var_param0, var_param1):
        a, b = _host(var_param0), _host(var_param1)
        d = self.dim
        mean_diff = a[:d] - b[:d]
        dl = a[d:] - b[d:]
        return .5 * np.sum(np.exp(2 * dl) + mean_diff ** 2 / np.exp(2 * b[d:]) - 2 * dl - 1)

    def mean_and_cov(self, var_param):
        vp = _host(var_param)
        return vp[:self.dim].copy(), np.diag(np.exp(2 * vp[self.dim:]))

    def _pth_moment(self, var_param, p):
        v = np.exp(2 * _host(var_param)[self.dim:])
        if p == 2:
            return np.sum(v)
        return 2 * np.sum(v ** 2) + np.sum(v) ** 2

    def supports_pth_moment(self, p):
        return p in [2, 4]


class MFStudentT(_MeanField):
    """A mean-field Student's t approximation family (approximations.py:254-312)."""
    _family = _lib.FAMILY_MF_STUDENT

    def __init__(self, dim, df, seed=1):
        if df <= 2:
            raise ValueError('df must be greater than 2')
        self._df = df
        super().__init__(dim, False, seed)

    def _draw(self, n, seed, offset):
        out = torch.empty(n, dtype=F64, device=device())
        _lib.check(_lib.lib.vb_philox_student_t_f64(_lib.ptr(out), n, float(self._df), seed, offset,
                                                     int(self.quantize_draws), _lib.stream()))
        return out

    def entropy(self, var_param):
        # df-only constants dropped, exactly as approximations.py:276-279
        return np.sum(_host(var_param)[self.dim:])

    def mean_and_cov(self, var_param):
        vp = _host(var_param)
        df = self.df
        return vp[:self.dim].copy(), df / (df - 2) * np.diag(np.exp(2 * vp[self.dim:]))

    def _pth_moment(self, var_param, p):
        df = self.df
        if df <= p:
            raise ValueError('df must be greater than p')
        s2 = np.exp(2 * _host(var_param)[self.dim:])
        c = df / (df - 2)
        if p == 2:
            return c * np.sum(s2)
        return c ** 2 * (2 * (df - 1) / (df - 4) * np.sum(s2 ** 2) + np.sum(s2) ** 2)

    def supports_pth_moment(self, p):
        return p in [2, 4] and p < self.df
