"""VI diagnostics -- drop-in mirror of viabel/diagnostics.py (all_diagnostics :13-64,
error_bounds :73-103, wasserstein_bounds :106-145, divergence_bound :148-186).

The O(n) reductions over the log weights (and over samples, when given) run on the device;
the scalar algebra on top of them is host Python.
"""
from warnings import warn

import numpy as np
import torch

from . import _lib
from ._tensor import F64, device, to_dev

__all__ = ['all_diagnostics', 'error_bounds', 'wasserstein_bounds', 'divergence_bound', 'sample_moments']


def _moments(log_weights, alpha):
    """max, sum r, sum lw, sum (lw-max)^2, sum r^2 with r = exp(lw-max)^alpha: one kernel pair, one D2H copy."""
    lw = to_dev(log_weights).reshape(-1)
    out = torch.empty(8, dtype=F64, device=lw.device)
    _lib.check(_lib.lib.vb_divergence_moments_f64(_lib.ptr(lw), lw.numel(), float(alpha), _lib.ptr(out),
                                                  _lib.stream()))
    h = out.cpu().numpy()
    return lw.numel(), float(h[0]), float(h[1]), float(h[2]), float(h[4]), float(h[5])


def _mc_warn(mean, second_moment, n, name, atol=0.01):
    """mean_and_check_mc_error (diagnostics.py:189-198): warn when std(a) / sqrt(n) > atol."""
    s = np.sqrt(max(second_moment - mean * mean, 0.0)) / np.sqrt(n)
    if s > atol:  # pragma: no cover
        warn('significant Monte Carlo error when computing {} (mean = {}, standard deviation = {})'
             .format(name, mean, s))


def divergence_bound(log_weights, *, alpha=2., log_norm_bound=None, return_log_norm_bound=False):
    """Bound on the alpha-divergence (diagnostics.py:148-186)."""
    if alpha <= 1:
        raise ValueError('alpha must be greater than 1')
    n, mx, sexp, sx, sc2, sexp2 = _moments(log_weights, alpha)
    _mc_warn(sexp / n, sexp2 / n, n, 'CUBO')
    cubo = np.log(sexp / n) / alpha + mx
    if log_norm_bound is None:
        log_norm_bound = sx / n
        _mc_warn(log_norm_bound - mx, sc2 / n, n, 'ELBO')      # moments of lw - max: same variance
    dalpha = alpha / (alpha - 1) * (cubo - log_norm_bound)
    if return_log_norm_bound:
        return dalpha, log_norm_bound
    return dalpha


def sample_moments(samples, want_cov=False):
    """(mean[d], m2[d], m4[d], cov[d,d] or None) of samples[n,d] through vb_sample_moments_f64:
    m2 / m4 are the per-coordinate central power sums sum_n (x_nj - mean_j)^{2,4}; cov = np.cov(samples.T)."""
    x = to_dev(samples)
    if x.dim() == 1:
        x = x[:, None].contiguous()
    n, d = int(x.shape[0]), int(x.shape[1])
    out = torch.empty(3 * d + (d * d if want_cov else 0), dtype=F64, device=x.device)
    mean, m2, m4 = out[:d], out[d:2 * d], out[2 * d:3 * d]
    cov = out[3 * d:].view(d, d) if want_cov else None
    ws = torch.empty(max(8, _lib.lib.vb_sample_moments_workspace_bytes(n, d, int(want_cov))), dtype=torch.uint8,
                     device=x.device)
    _lib.check(_lib.lib.vb_sample_moments_f64(_lib.ptr(x), n, d, x.stride(0), _lib.ptr(mean), _lib.ptr(m2), _lib.ptr(m4),
                                              _lib.ptr(cov), _lib.ptr(ws), ws.numel(), _lib.stream()))
    return mean, m2, m4, cov


def wasserstein_bounds(d2, *, samples=None, moment_bound_fn=None):
    """1- and 2-Wasserstein bounds from a 2-divergence bound (diagnostics.py:106-145)."""
    results = dict()
    if moment_bound_fn is None:
        if samples is None:
            raise ValueError('must provides samples if moment_bound_fn not given')
        n = int(np.shape(samples)[0]) if not isinstance(samples, torch.Tensor) else int(samples.shape[0])
        _, m2, m4, _ = sample_moments(samples)
        sums = {2: float(m2.sum()), 4: float(m4.sum())}

        def moment_bound_fn(p):
            # mean over draws of the per-coordinate central power sums, as the reference (diagnostics.py:140-141)
            return sums[p] / n
    for p in [1, 2]:
        Cp = moment_bound_fn(2 * p)
        results['W{}'.format(p)] = 2 * Cp ** (.5 / p) * np.expm1(d2) ** (.5 / p)
    return results


def _compute_norm_if_needed(var):
    if isinstance(var, torch.Tensor):
        var = var.detach().cpu().numpy()
    if np.asarray(var).ndim == 2:
        return np.linalg.norm(var, ord=2)
    return var


def error_bounds(*, W1=np.inf, W2=np.inf, q_var=np.inf, p_var=np.inf):
    """Mean / std / covariance error bounds (diagnostics.py:73-103, :201-219)."""
    qv, pv = _compute_norm_if_needed(q_var), _compute_norm_if_needed(p_var)
    min_var = qv if pv is None else np.min([qv, pv], axis=0)
    return dict(mean_error=min(W1, W2), std_error=W2,
                cov_error=2 * (np.sqrt(min_var) * W2 + W2 ** 2))


def all_diagnostics(log_weights, *, samples=None, moment_bound_fn=None, q_var=None, p_var=None,
                    log_norm_bound=None):
    """All VI diagnostics (diagnostics.py:13-64)."""
    d2, log_norm_bound = divergence_bound(log_weights, log_norm_bound=log_norm_bound,
                                          return_log_norm_bound=True)
    results = wasserstein_bounds(d2, samples=samples, moment_bound_fn=moment_bound_fn)
    if q_var is None and samples is not None:
        q_var = sample_moments(samples, want_cov=True)[3].cpu().numpy()   # np.cov(samples.T), SYRK on the FP64 tensor pipe
        if q_var.shape == (1, 1):
            q_var = q_var.reshape(())                                      # np.cov of one variable is a scalar
    results.update(error_bounds(q_var=q_var, p_var=p_var, **results))
    results['d2'] = d2
    results['log_norm_bound'] = log_norm_bound
    return results
