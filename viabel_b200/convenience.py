"""One-call entry points with viabel's signatures (bbvi: convenience.py:14-94, vi_diagnostics: :97-179).

Everything here is argument plumbing around the device path: pick the objective and the optimiser,
run it, then draw -> log-weights -> PSIS -> bounds for the diagnostics.  Messages and exceptions are
the reference's, so that callers and tests written against viabel keep working."""
from ._psis import psislw
from ._tensor import is_host, to_dev
from .approximations import MFGaussian
from .diagnostics import all_diagnostics
from .models import Model
from .objectives import ExclusiveKL
from .optimization import RMSProp

__all__ = ['bbvi', 'vi_diagnostics', 'psis_correction', 'samples_and_log_weights']

_KHAT_LIMIT = 0.7          # importance sampling is hopeless above this (convenience.py:143)
_D2_BAD, _D2_OK = 4.6, 0.1  # 2-divergence bands of the verdict printed at the end (:160-166)


def _objective_for(dimension, num_mc_samples, log_density, approx, objective, fit):
    """Resolve the (objective, approx) pair from the mutually exclusive ways of specifying a problem."""
    if objective is not None:
        if not (fit is None and log_density is None and approx is None):
            raise ValueError('if objective is specified, cannot specify fit, log_density, or approx')
        return objective, objective.approx
    if log_density is not None and fit is not None:
        raise ValueError('log_density and fit cannot both be specified')
    if log_density is None:
        if fit is None:
            raise ValueError('either log_density or fit must be specified if objective not given')
        raise NotImplementedError('StanModel is out of scope of the B200 hot path')
    model = log_density if isinstance(log_density, Model) else Model(log_density)
    family = MFGaussian(dimension) if approx is None else approx
    return ExclusiveKL(family, model, num_mc_samples), family


def _optimizer_for(adaptive, fixed_lr, learning_rate, RMS_kwargs, FASO_kwargs, RAABBVI_kwargs):
    """(adaptive, fixed_lr) -> RAABBVI | FASO | plain RMSProp, all driving the same RMSProp direction."""
    from .optimization import FASO, RAABBVI
    if not adaptive and not fixed_lr:
        raise ValueError('if fixed_lr is False, adaptive must be True')
    inner = RMSProp(learning_rate, diagnostics=True, **RMS_kwargs)
    if not adaptive:
        return inner
    return FASO(inner, **FASO_kwargs) if fixed_lr else RAABBVI(inner, **RAABBVI_kwargs)


def bbvi(dimension, *, n_iters=10000, num_mc_samples=10, log_density=None, approx=None, objective=None,
         fit=None, adaptive=True, fixed_lr=False, init_var_param=None, learning_rate=0.01,
         RMS_kwargs=dict(), FASO_kwargs=dict(), RAABBVI_kwargs=dict()):
    """Black-box variational inference in one call.  `log_density` is a callable on CUDA tensors or a
    Model plugin (e.g. LogisticRegression); the result dictionary is the optimiser's plus `objective`."""
    objective, family = _objective_for(dimension, num_mc_samples, log_density, approx, objective, fit)
    optimizer = _optimizer_for(adaptive, fixed_lr, learning_rate, RMS_kwargs, FASO_kwargs, RAABBVI_kwargs)
    start = family.init_param() if init_var_param is None else init_var_param
    results = optimizer.optimize(n_iters, objective, start)
    results['objective'] = objective
    return results


def vi_diagnostics(var_param, *, objective=None, model=None, approx=None, n_samples=100000):
    """Pareto k-hat, 2-divergence and the error bounds for a fitted approximation."""
    if objective is not None:
        if model is not None or approx is not None:
            raise ValueError('model and/or approx cannot be specified if objective is')
        model, approx = objective.model, objective.approx
    elif model is None or approx is None:
        raise ValueError('either objective or both model and approx must be specified')
    if n_samples <= 0:
        raise ValueError('n_samples must be positive')
    return _vi_diagnostics(var_param, model, approx, n_samples)


def _d2_verdict(d2):
    if d2 > _D2_BAD:  # pragma: no cover
        return 'WARNING: d2 > 4.6 means the approximation is very inaccurate'
    if d2 > _D2_OK:
        return ('WARNING: 0.1 < d2 < 4.6 means the approximation is somewhat '
                'inaccurate. Use importance sampling to decrease error.')
    return '\nAll diagnostics pass.'


def _vi_diagnostics(var_param, model, approx, n_samples, base=None):
    """draw -> log weights -> PSIS; stop at a hopeless k-hat, else add the divergence / Wasserstein bounds."""
    samples, slw, khat = psis_correction(var_param, model, approx, n_samples, base=base)
    report = {'samples': samples, 'smoothed_log_weights': slw, 'khat': khat}
    print('Pareto k is estimated to be khat = {:.2f}'.format(khat))
    if khat > _KHAT_LIMIT:
        print('WARNING: khat > 0.7 means importance sampling is not feasible.')
        print('WARNING: not running further diagnostics')
        return report
    print()
    closed_form = approx.supports_pth_moment(2) and approx.supports_pth_moment(4)
    moments = (lambda p: approx.pth_moment(var_param, p)) if closed_form else None
    report.update(all_diagnostics(slw, samples=samples, moment_bound_fn=moments,
                                  q_var=approx.mean_and_cov(var_param)[1]))
    print('The 2-divergence is estimated to be d2 = {:.2g}'.format(report['d2']))
    print(_d2_verdict(report['d2']))
    return report


def psis_correction(var_param, model, approx, n_samples, base=None):
    """(samples [dim, n] -- transposed like the reference --, smoothed log weights, k-hat)."""
    draws, lw = samples_and_log_weights(var_param, model, approx, n_samples, base=base)
    slw, khat = psislw(lw, overwrite_lw=True)
    return draws.T, slw, khat


def samples_and_log_weights(var_param, model, approx, n_samples, base=None):
    """n draws from the approximation and log p - log q at them; numpy in -> numpy out."""
    target = model if isinstance(model, Model) else Model(model)
    vp = to_dev(var_param)
    draws = approx.sample(vp, n_samples, **({} if base is None else {'base': base}))
    lw = target(draws) - approx.log_density(vp, draws)
    if is_host(var_param):
        return draws.cpu().numpy(), lw.cpu().numpy()
    return draws, lw
