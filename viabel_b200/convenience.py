"""Convenience entry points -- mirror of viabel/convenience.py (bbvi :14-94,
vi_diagnostics :97-179)."""
import numpy as np
import torch

from ._psis import psislw
from ._tensor import is_host, to_dev
from .approximations import MFGaussian
from .diagnostics import all_diagnostics
from .models import Model
from .objectives import ExclusiveKL
from .optimization import RMSProp

__all__ = ['bbvi', 'vi_diagnostics', 'psis_correction', 'samples_and_log_weights']


def bbvi(dimension, *, n_iters=10000, num_mc_samples=10, log_density=None, approx=None, objective=None,
         fit=None, adaptive=True, fixed_lr=False, init_var_param=None, learning_rate=0.01,
         RMS_kwargs=dict(), FASO_kwargs=dict(), RAABBVI_kwargs=dict()):
    """Fit a model using black-box variational inference (convenience.py:14-94).

    `log_density` is a callable on CUDA tensors or a Model plugin."""
    from .optimization import FASO, RAABBVI
    if objective is not None:
        if fit is not None or log_density is not None or approx is not None:
            raise ValueError('if objective is specified, cannot specify fit, log_density, or approx')
        approx = objective.approx
    else:
        if log_density is None:
            if fit is None:
                raise ValueError('either log_density or fit must be specified if objective not given')
            raise NotImplementedError('StanModel is out of scope of the B200 hot path')
        elif fit is None:
            model = log_density if isinstance(log_density, Model) else Model(log_density)
        else:
            raise ValueError('log_density and fit cannot both be specified')
        if approx is None:
            approx = MFGaussian(dimension)
        objective = ExclusiveKL(approx, model, num_mc_samples)
    if init_var_param is None:
        init_var_param = approx.init_param()
    base_opt = RMSProp(learning_rate, diagnostics=True, **RMS_kwargs)
    if adaptive and not fixed_lr:
        opt = RAABBVI(base_opt, **RAABBVI_kwargs)
    elif adaptive and fixed_lr:
        opt = FASO(base_opt, **FASO_kwargs)
    elif not adaptive and fixed_lr:
        opt = base_opt
    else:
        raise ValueError('if fixed_lr is False, adaptive must be True')
    opt_results = opt.optimize(n_iters, objective, init_var_param)
    opt_results['objective'] = objective
    return opt_results


def vi_diagnostics(var_param, *, objective=None, model=None, approx=None, n_samples=100000):
    """Pareto k-hat and 2-divergence diagnostics with error bounds (convenience.py:97-133)."""
    if objective is None:
        if model is None or approx is None:
            raise ValueError('either objective or both model and approx must be specified')
    elif model is not None or approx is not None:
        raise ValueError('model and/or approx cannot be specified if objective is')
    else:
        model = objective.model
        approx = objective.approx
    if n_samples <= 0:
        raise ValueError('n_samples must be positive')
    return _vi_diagnostics(var_param, model, approx, n_samples)


def _vi_diagnostics(var_param, model, approx, n_samples, base=None):
    host = is_host(var_param)
    samples, smoothed_log_weights, khat = psis_correction(var_param, model, approx, n_samples, base=base)
    results = dict(samples=samples, smoothed_log_weights=smoothed_log_weights, khat=khat)
    print('Pareto k is estimated to be khat = {:.2f}'.format(results['khat']))
    if results['khat'] > 0.7:
        print('WARNING: khat > 0.7 means importance sampling is not feasible.')
        print('WARNING: not running further diagnostics')
        return results
    print()
    if approx.supports_pth_moment(2) and approx.supports_pth_moment(4):
        def moment_bound_fn(p):
            return approx.pth_moment(var_param, p)
    else:
        moment_bound_fn = None
    _, q_var = approx.mean_and_cov(var_param)
    results.update(all_diagnostics(smoothed_log_weights, samples=samples, moment_bound_fn=moment_bound_fn,
                                   q_var=q_var))
    print('The 2-divergence is estimated to be d2 = {:.2g}'.format(results['d2']))
    if results['d2'] > 4.6:  # pragma: no cover
        print('WARNING: d2 > 4.6 means the approximation is very inaccurate')
    elif results['d2'] > 0.1:
        print('WARNING: 0.1 < d2 < 4.6 means the approximation is somewhat '
              'inaccurate. Use importance sampling to decrease error.')
    else:
        print('\nAll diagnostics pass.')
    return results


def psis_correction(var_param, model, approx, n_samples, base=None):
    """convenience.py:170-173 (returns samples transposed, [dim, n])."""
    samples, log_weights = samples_and_log_weights(var_param, model, approx, n_samples, base=base)
    smoothed_log_weights, khat = psislw(log_weights, overwrite_lw=True)
    return samples.T, smoothed_log_weights, khat


def samples_and_log_weights(var_param, model, approx, n_samples, base=None):
    """convenience.py:176-179"""
    if not isinstance(model, Model):
        model = Model(model)
    host = is_host(var_param)
    vp = to_dev(var_param)
    samples = approx.sample(vp, n_samples) if base is None else approx.sample(vp, n_samples, base=base)
    log_weights = model(samples) - approx.log_density(vp, samples)
    if host:
        return samples.cpu().numpy(), log_weights.cpu().numpy()
    return samples, log_weights
