"""NeuralNet and NVPFlow families -- mirror of viabel/approximations.py:385-550.

These are the reference's generic-AD families: small coupling networks whose parameters are the variational
parameters.  Here the networks are evaluated on the device in float64 with torch ops and differentiated by torch
autograd (the same arrangement as LRGaussian: the model's log density and gradient inside an objective still come
from the model plugin's own kernels).  They are off the north-star hot path -- O(n * sum of layer sizes) work -- and
exist so that every family of the reference has a drop-in.

Layouts follow paragami's PatternDict (insertion order, C-order ravel):
  NeuralNet  : ["0" (W0), "0_b" (b0), "1", "1_b", ...]
  NVPFlow    : ["0t" (NeuralNet t_0), "0s" (NeuralNet s_0), "1t", "1s", ...]
"""
import numpy as np
import torch

from ._tensor import F64, is_host, like_input, to_dev
from .approximations import ApproximationFamily, MFGaussian, MFStudentT, _MeanField

__all__ = ['NeuralNet', 'NVPFlow']

_PROBE = np.array([-1.3, -0.2, 0.0, 0.4, 2.1])


def _relu(x):
    return torch.clamp(x, min=0.0)


def _torch_activation(fn):
    """(torch function, its elementwise derivative or None) for a numpy / torch callable.  The reference takes numpy
    callables (np.tanh, lambda x: x, ...); they are recognised by their values on a probe vector."""
    if fn is None:
        return (lambda x: x), (lambda x: torch.ones_like(x))
    if isinstance(fn, str):
        fn = {'tanh': np.tanh, 'identity': (lambda x: x), 'relu': (lambda x: np.maximum(x, 0.0)),
              'sigmoid': (lambda x: 1.0 / (1.0 + np.exp(-x)))}[fn]
    try:
        vals = np.asarray(fn(_PROBE.copy()), dtype=np.float64)
    except Exception:
        vals = None
    if vals is not None and vals.shape == _PROBE.shape:
        if np.allclose(vals, np.tanh(_PROBE), rtol=0, atol=1e-15):
            return torch.tanh, (lambda x: 1.0 - torch.tanh(x) ** 2)
        if np.array_equal(vals, _PROBE):
            return (lambda x: x), (lambda x: torch.ones_like(x))
        if np.array_equal(vals, np.maximum(_PROBE, 0.0)):
            return _relu, (lambda x: (x > 0).to(x.dtype))
        if np.allclose(vals, 1.0 / (1.0 + np.exp(-_PROBE)), rtol=0, atol=1e-15):
            return torch.sigmoid, (lambda x: torch.sigmoid(x) * (1.0 - torch.sigmoid(x)))
    # a callable that works on tensors: differentiate it with autograd (elementwise_grad, approximations.py:416-417)
    try:
        fn(torch.zeros(2, dtype=F64))
    except Exception:
        raise NotImplementedError('activation must be tanh / identity / relu / sigmoid or a torch-callable function')
    return fn, None


def _elementwise_grad(fn, dfn, x):
    if dfn is not None:
        return dfn(x)
    with torch.enable_grad():
        xx = x if x.requires_grad else x.detach().requires_grad_(True)
        (g,) = torch.autograd.grad(fn(xx).sum(), xx, create_graph=True)
    return g


class NeuralNet(ApproximationFamily):
    """approximations.py:385-449.  `var_param` of forward / sample is the FOLDED parameter dict ({"0": W0, "0_b": b0,
    ...}), as in the reference (its tests fold before calling, tests/test_approximations.py:122-124); a flat vector
    in PatternDict order is accepted as well."""

    def __init__(self, layers_shapes, nonlinearity=np.tanh, last=np.tanh, mc_samples=10000, seed=1):
        self.mc_samples = mc_samples
        self._shapes = [tuple(int(v) for v in s) for s in layers_shapes]
        self._layers = len(self._shapes)
        self._nonlinearity, self._dnonlinearity = _torch_activation(nonlinearity)
        self._last, self._dlast = _torch_activation(last)
        self._seed = int(seed)
        self.input_dim = self._shapes[0][0]
        n = sum(a * b + b for a, b in self._shapes)
        super().__init__(self._shapes[-1][-1], n, False, False)

    # -- parameter handling -------------------------------------------------------------------
    def fold(self, flat):
        """flat vector -> {"0": W0, "0_b": b0, ...} (views of the flat tensor: autograd flows through)."""
        flat = flat.reshape(-1)
        if flat.numel() != self.var_param_dim:
            raise ValueError('var_param has the wrong length')
        out, off = {}, 0
        for i, (a, b) in enumerate(self._shapes):
            out[str(i)] = flat[off:off + a * b].reshape(a, b)
            off += a * b
            out[str(i) + '_b'] = flat[off:off + b]
            off += b
        return out

    def _params(self, var_param):
        if isinstance(var_param, dict):
            return {k: to_dev(v) if not isinstance(v, torch.Tensor) else v for k, v in var_param.items()}
        return self.fold(var_param if isinstance(var_param, torch.Tensor) else to_dev(var_param))

    def forward_t(self, params, x):
        """(output, log_det_J) on tensors -- the reference's formula verbatim (:414-429): the 'log determinant' term
        is log|sum_j (f'(out) W^T)_j| with the derivative evaluated at the layer OUTPUT."""
        log_det_J = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
        for i in range(self._layers):
            W, b = params[str(i)], params[str(i) + '_b']
            lastl = i + 1 == self._layers
            f, df = (self._last, self._dlast) if lastl else (self._nonlinearity, self._dnonlinearity)
            x = f(x @ W + b)
            log_det_J = log_det_J + torch.log(torch.abs((_elementwise_grad(f, df, x) @ W.T).sum(dim=1)))
        return x, log_det_J

    def forward(self, var_param, x):
        host = is_host(x)
        xd = to_dev(x)
        y, ld = self.forward_t(self._params(var_param), xd)
        return like_input(y, host), like_input(ld, host)

    def sample(self, var_param, n_samples, seed=None, base=None):
        host = not isinstance(var_param, torch.Tensor) and not (
            isinstance(var_param, dict) and any(isinstance(v, torch.Tensor) for v in var_param.values()))
        if base is None:
            gen = torch.Generator(device='cuda')
            gen.manual_seed(self._seed if seed is None else int(seed))
            if seed is None:
                self._seed += 1
            base = torch.randn(int(n_samples), self.input_dim, generator=gen, device='cuda', dtype=F64)
        z0 = to_dev(base)
        self.last_base = z0
        return like_input(self.forward_t(self._params(var_param), z0)[0], host)

    def log_density(self, var_param, x):
        raise NotImplementedError

    def mean_and_cov(self, var_param):
        s = self.sample(var_param, self.mc_samples)
        s = s if isinstance(s, torch.Tensor) else to_dev(s)
        from .diagnostics import sample_moments
        mean, _, _, cov = sample_moments(s, want_cov=True)
        return mean.cpu().numpy().copy(), cov.cpu().numpy().copy()

    def _pth_moment(self, var_param, p):
        raise NotImplementedError

    def supports_pth_moment(self, p):
        return False


def prior_log_density_t(prior, prior_param, z):
    """Differentiable (torch) log density of a mean-field prior at z[n,d] (approximations.py:231-236, :281-286)."""
    if not isinstance(prior, _MeanField):
        raise NotImplementedError('NVPFlow priors: MFGaussian or MFStudentT')
    d = prior.dim
    mu, ls = prior_param[:d], prior_param[d:]
    u = (z - mu) * torch.exp(-ls)
    if isinstance(prior, MFGaussian):
        return (-0.5 * u * u - ls - 0.5 * np.log(2 * np.pi)).sum(dim=1)
    df = float(prior.df)
    c = float(torch.lgamma(torch.tensor((df + 1) / 2, dtype=F64)) - torch.lgamma(torch.tensor(df / 2, dtype=F64))) \
        - 0.5 * np.log(df * np.pi)
    return (c - 0.5 * (df + 1) * torch.log1p(u * u / df) - ls).sum(dim=1)


class NVPFlow(ApproximationFamily):
    """Real NVP flow over a mean-field prior (approximations.py:452-550).  No entropy / KL / moments: objectives use
    the path-derivative ExclusiveKL or AlphaDivergence on it (objectives._flow_objective)."""

    def __init__(self, layers_t, layers_s, mask, prior, prior_param, dim, activation=np.tanh, seed=1, mc_samples=10000):
        assert len(layers_t) == len(layers_s)
        self.prior = prior
        self.prior_param = np.asarray(prior_param, dtype=np.float64)
        self.mc_samples = mc_samples
        self._dim = int(dim)
        self._seed = int(seed)
        self.mask = np.asarray(mask, dtype=np.float64)
        self.t = [NeuralNet(layers_t, nonlinearity=activation, last=lambda x: x) for _ in range(len(self.mask))]
        self.s = [NeuralNet(layers_s, nonlinearity=activation, last=np.tanh) for _ in range(len(self.mask))]
        n = sum(t.var_param_dim + s.var_param_dim for t, s in zip(self.t, self.s))
        self.last_base = None
        super().__init__(self._dim, n, False, False)

    def _fold(self, vp):
        """flat -> [(t_params, s_params)] per coupling layer, "it" before "is" (:487-489)."""
        vp = vp.reshape(-1)
        if vp.numel() != self.var_param_dim:
            raise ValueError('var_param has the wrong length')
        out, off = [], 0
        for t, s in zip(self.t, self.s):
            tp = t.fold(vp[off:off + t.var_param_dim])
            off += t.var_param_dim
            sp = s.fold(vp[off:off + s.var_param_dim])
            off += s.var_param_dim
            out.append((tp, sp))
        return out

    def _masks(self, ref):
        return torch.as_tensor(self.mask, dtype=ref.dtype, device=ref.device)

    def g_t(self, vp, z):
        """Latent -> data space (:493-511)."""
        params, masks = self._fold(vp), self._masks(z)
        x = z
        for i, (tp, sp) in enumerate(params):
            m = masks[i]
            x_ = x * m
            s = self.s[i].forward_t(sp, x_)[0] * (1 - m)
            t = self.t[i].forward_t(tp, x_)[0] * (1 - m)
            x = x_ + (1 - m) * (x * torch.exp(s) + t)
        return x

    def f_t(self, vp, x):
        """Data -> latent space and the log determinant (:513-531)."""
        params, masks = self._fold(vp), self._masks(x)
        log_det_J = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
        z = x
        for i in reversed(range(len(params))):
            tp, sp = params[i]
            m = masks[i]
            z_ = m * z
            s = self.s[i].forward_t(sp, z_)[0] * (1 - m)
            t = self.t[i].forward_t(tp, z_)[0] * (1 - m)
            z = (1 - m) * (z - t) * torch.exp(-s) + z_
            log_det_J = log_det_J - s.sum(dim=1)
        return z, log_det_J

    def log_density_t(self, vp, x):
        z, logdet = self.f_t(vp, x)
        pp = torch.as_tensor(self.prior_param, dtype=x.dtype, device=x.device)
        return prior_log_density_t(self.prior, pp, z) + logdet

    # -- public API (numpy in -> numpy out, tensors in -> tensors out) ------------------------
    def g(self, var_param, z):
        host = is_host(z)
        return like_input(self.g_t(to_dev(var_param), to_dev(z)), host)

    def f(self, var_param, x):
        host = is_host(x)
        z, ld = self.f_t(to_dev(var_param), to_dev(x))
        return like_input(z, host), like_input(ld, host)

    def log_density(self, var_param, x):
        host = is_host(x)
        xd = to_dev(x)
        if xd.dim() == 1:
            xd = xd[None, :]
        return like_input(self.log_density_t(to_dev(var_param), xd), host)

    def prior_draws(self, n_samples, seed=None, base=None):
        """z_0 = prior.sample(prior_param, n, seed) (:537-538) on the device; `base` injects the prior's base draws."""
        z0 = self.prior.sample(to_dev(self.prior_param), int(n_samples), seed=seed, base=base)
        self.last_base = self.prior.last_base
        return z0

    def sample(self, var_param, n_samples, seed=None, base=None):
        host = is_host(var_param)
        z0 = self.prior_draws(n_samples, seed, base)
        return like_input(self.g_t(to_dev(var_param), z0), host)

    def mean_and_cov(self, var_param):
        s = self.sample(to_dev(var_param), self.mc_samples)
        from .diagnostics import sample_moments
        mean, _, _, cov = sample_moments(s.contiguous(), want_cov=True)
        return mean.cpu().numpy().copy(), cov.cpu().numpy().copy()

    def _pth_moment(self, var_param, p):
        raise NotImplementedError

    def supports_pth_moment(self, p):
        return False
