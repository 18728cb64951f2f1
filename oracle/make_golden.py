"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference's viabel/*.py files are imported as they lie under /root/reference;
`autograd`, `paragami` and `pystan` (absent from the image) are replaced by the
stand-ins in oracle/refshim (torch float64 provides the reverse-mode AD that autograd
would).  viabel/_psis.py and viabel/diagnostics.py need numpy only and run untouched.

Every RandomState draw the reference makes is recorded, so the goldens carry the
base draws (eps / t / (chi2, z)) next to the outputs: parity is by draw injection.
"""
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(1, '/root/reference')

import numpy as np  # noqa: E402

import autograd.numpy as anp  # noqa: E402
import autograd.numpy.random as shim_random  # noqa: E402
from autograd.scipy.stats import norm as anorm  # noqa: E402
from autograd.scipy.stats import t as at_dist  # noqa: E402
from autograd.scipy.special import gammaln as agammaln  # noqa: E402,F401

import viabel  # noqa: E402
from viabel import _psis as ref_psis  # noqa: E402
from viabel import diagnostics as ref_diag  # noqa: E402
from viabel.approximations import MFGaussian, MFStudentT, MultivariateT  # noqa: E402
from viabel.objectives import AlphaDivergence, ExclusiveKL  # noqa: E402
from viabel.optimization import Adam, RMSProp  # noqa: E402

OUT = os.path.join(HERE, '..', 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


# ---------------------------------------------------------------------------
# synthetic problems (shared with tests/_problems.py, which rebuilds them by seed)
# ---------------------------------------------------------------------------
sys.path.insert(0, os.path.join(HERE, '..', 'tests'))
from _problems import (  # noqa: E402
    diag_problem, hier_problem, logistic_problem, psis_case, PSIS_CASES, target_params)


def logistic_log_p(X, y, prior_sd):
    def f(theta):
        z = anp.dot(theta, X.T) * y
        d = X.shape[1]
        return (-anp.sum(anp.logaddexp(0.0, -z), axis=1)
                - 0.5 * anp.sum(theta ** 2, axis=1) / prior_sd ** 2
                - d * math.log(prior_sd * math.sqrt(2 * math.pi)))
    return f


def probit_log_p(X, y, prior_sd):
    def f(theta):
        z = anp.dot(theta, X.T) * y
        d = X.shape[1]
        return (anp.sum(anorm.logcdf(z), axis=1)
                - 0.5 * anp.sum(theta ** 2, axis=1) / prior_sd ** 2
                - d * math.log(prior_sd * math.sqrt(2 * math.pi)))
    return f


def gauss_log_p(mean, sd):
    return lambda x: anp.sum(anorm.logpdf(x, loc=mean, scale=sd), axis=1)


def student_log_p(loc, scale, df):
    return lambda x: anp.sum(at_dist.logpdf(x, df, loc, scale), axis=1)


def hier_log_p(X, y, group, G, p):
    N = X.shape[0]
    onehot = np.zeros((N, G))
    onehot[np.arange(N), group] = 1.0

    def f(theta):
        S = theta.shape[0]
        beta = theta[:, :G * p].reshape((S, G, p))
        m = theta[:, G * p:G * p + p]
        ltau, lsig = theta[:, -2], theta[:, -1]
        lp = 0.0
        for g in range(G):
            sel = group == g
            pred = anp.dot(beta[:, g, :], X[sel].T)                 # [S, n_g]
            lp = lp + anp.sum(anorm.logpdf(y[sel], pred, anp.exp(lsig)[:, None]), axis=1)
            lp = lp + anp.sum(anorm.logpdf(beta[:, g, :], m, anp.exp(ltau)[:, None]), axis=1)
        lp = lp + anp.sum(anorm.logpdf(m, 0.0, 10.0), axis=1)
        lp = lp + anorm.logpdf(ltau, 0.0, 1.0) + anorm.logpdf(lsig, 0.0, 1.0)
        return lp
    return f


class Recorder(object):
    def __enter__(self):
        shim_random.RECORD = []
        return shim_random.RECORD

    def __exit__(self, *a):
        shim_random.RECORD = None


def first_draws(rec, family):
    """Base draws of the FIRST sample() call in a recording."""
    if family == 'mvt':
        assert rec[0][0] == 'chisquare' and rec[1][0] == 'randn'
        return dict(chi2=rec[0][2], z=rec[1][2])
    assert rec[0][0] in ('randn', 'standard_t')
    return dict(base=rec[0][2])


def make_family(kind, d, df, seed):
    if kind == 'mfg':
        return MFGaussian(d, seed=seed)
    if kind == 'mft':
        return MFStudentT(d, df, seed=seed)
    return MultivariateT(d, df, seed=seed)


def random_var_param(fam, kind, d, rs):
    if kind == 'mvt':
        B = rs.randn(d, d) * 0.4 + np.eye(d)
        Sigma = B @ B.T + 0.3 * np.eye(d)
        return fam._pattern.flatten(dict(mu=rs.randn(d), Sigma=Sigma))
    return np.concatenate([rs.randn(d), 0.5 * rs.randn(d) - 0.5])


# ---------------------------------------------------------------------------
def gen_families():
    out = {}
    rs = np.random.RandomState(341)
    for kind, df in (('mfg', None), ('mft', 20), ('mft', 5.5), ('mvt', 100), ('mvt', 7)):
        for d in (1, 3, 8):
            tag = '%s_df%s_d%d' % (kind, df, d)
            fam = make_family(kind, d, df, seed=226)
            vp = random_var_param(fam, kind, d, rs)
            vp1 = random_var_param(fam, kind, d, rs)
            with Recorder() as rec:
                x = fam.sample(vp, 64)
            for k, v in first_draws(rec, kind).items():
                out[tag + '/' + k] = v
            out[tag + '/var_param'] = vp
            out[tag + '/var_param1'] = vp1
            out[tag + '/init_param'] = fam.init_param()
            out[tag + '/sample'] = x
            out[tag + '/log_density'] = np.asarray(fam.log_density(vp, x))
            out[tag + '/log_density_1d'] = np.asarray(fam.log_density(vp, x[0]))
            out[tag + '/entropy'] = fam.entropy(vp)
            if fam.supports_kl:
                out[tag + '/kl'] = fam.kl(vp, vp1)
            mean, cov = fam.mean_and_cov(vp)
            out[tag + '/mean'] = mean
            out[tag + '/cov'] = cov
            for p in (2, 4):
                if fam.supports_pth_moment(p):
                    out[tag + '/moment%d' % p] = fam.pth_moment(vp, p)
    np.savez_compressed(os.path.join(OUT, 'families.npz'), **out)
    print('families: %d arrays' % len(out))


def gen_objectives():
    out = {}
    rs = np.random.RandomState(851)
    models = {}
    X, y, _ = logistic_problem(60, 4, seed=11)
    models['logistic_d4'] = (4, logistic_log_p(X, y, 10.0))
    X, y, _ = logistic_problem(1000, 10, seed=12)
    models['logistic_d10'] = (10, logistic_log_p(X, y, 10.0))
    X, y, _ = logistic_problem(200, 6, seed=13)
    models['probit_d6'] = (6, probit_log_p(X, y, 10.0))
    mean, sd = target_params(5, seed=14)
    models['gauss_d5'] = (5, gauss_log_p(mean, sd))
    models['student_d5'] = (5, student_log_p(mean, sd, 10.0))
    hp = hier_problem(G=3, p=2, n_per=7, seed=15)
    models['hier_G3p2'] = (3 * 2 + 2 + 2, hier_log_p(hp['X'], hp['y'], hp['group'], 3, 2))

    for mname, (d, logp) in models.items():
        for kind, df in (('mfg', None), ('mft', 8), ('mvt', 9)):
            fam = make_family(kind, d, df, seed=1214)
            for point in ('init', 'rand'):
                if point == 'init' and kind == 'mvt':
                    continue        # Sigma = 10 I is degenerate for eigh's VJP
                vp = fam.init_param() if point == 'init' else random_var_param(fam, kind, d, rs)
                if mname.startswith('hier') and point == 'init':
                    continue        # exp(log_sigma=2) draws overflow the likelihood scale
                for S in (7,):
                    objs = {'ekl': lambda: ExclusiveKL(fam, logp, S),
                            'alpha2': lambda: AlphaDivergence(fam, logp, S, 2.0),
                            'alpha1.5': lambda: AlphaDivergence(fam, logp, S, 1.5)}
                    if kind != 'mvt':
                        objs['ekl_path'] = lambda: ExclusiveKL(fam, logp, S, use_path_deriv=True)
                    for oname, mk in objs.items():
                        tag = '%s/%s_df%s/%s/%s' % (mname, kind, df, point, oname)
                        obj = mk()
                        np.random.seed(5039)
                        with Recorder() as rec:
                            value, grad = obj(vp)
                        for k, v in first_draws(rec, kind).items():
                            out[tag + '/' + k] = v
                        out[tag + '/var_param'] = vp
                        out[tag + '/value'] = float(value)
                        out[tag + '/grad'] = grad
    np.savez_compressed(os.path.join(OUT, 'objectives.npz'), **out)
    print('objectives: %d arrays' % len(out))


def gen_objectives_cv():
    """ExclusiveKL with the control-variate estimators (objectives.py:170-273): every hessian_approx_method, with and
    without the path derivative, mean-field families.  autograd's hessian / make_hvp / elementwise_grad are the torch
    stand-ins of oracle/refshim."""
    out = {}
    rs = np.random.RandomState(1639)
    models = {}
    X, y, _ = logistic_problem(60, 4, seed=11)
    models['logistic_d4'] = (4, logistic_log_p(X, y, 10.0))
    X, y, _ = logistic_problem(1000, 10, seed=12)
    models['logistic_d10'] = (10, logistic_log_p(X, y, 10.0))
    X, y, _ = logistic_problem(200, 6, seed=13)
    models['probit_d6'] = (6, probit_log_p(X, y, 10.0))
    mean, sd = target_params(5, seed=14)
    models['gauss_d5'] = (5, gauss_log_p(mean, sd))
    models['student_d5'] = (5, student_log_p(mean, sd, 10.0))
    for mname, (d, logp2d) in models.items():
        # the estimators also evaluate the model at the 1-D variational mean (objectives.py:203, :221): the
        # reference's own test density broadcasts over that (tests/test_objectives.py:18-19); these promote it
        logp = (lambda f: (lambda th: f(anp.atleast_2d(th))))(logp2d)
        for kind, df in (('mfg', None), ('mft', 8)):
            fam = make_family(kind, d, df, seed=1214)
            vp = random_var_param(fam, kind, d, rs)
            for method in ('full', 'mean_only', 'loo_diag_approx', 'loo_direct_approx'):
                for path in (False, True):
                    tag = '%s/%s_df%s/%s/%s' % (mname, kind, df, method, 'path' if path else 'plain')
                    obj = ExclusiveKL(fam, logp, 9, use_path_deriv=path, hessian_approx_method=method)
                    with Recorder() as rec:
                        value, grad = obj(vp)
                    for k, v in first_draws(rec, kind).items():
                        out[tag + '/' + k] = v
                    out[tag + '/var_param'] = vp
                    out[tag + '/value'] = float(value)
                    out[tag + '/grad'] = np.asarray(grad, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, 'objectives_cv.npz'), **out)
    print('objectives_cv: %d arrays' % len(out))


def gen_lr_gaussian():
    """LRGaussian (approximations.py:610-731): family functions, then ExclusiveKL / AlphaDivergence on it."""
    from viabel.approximations import LRGaussian
    out = {}
    rs = np.random.RandomState(153)
    for d, k in ((3, 0), (3, 1), (8, 3), (6, 6)):
        tag = 'lr_d%d_k%d' % (d, k)
        fam = LRGaussian(d, seed=226, k=k)
        vp = np.concatenate([rs.randn(d), 0.4 * rs.randn(d) - 0.3, 0.5 * rs.randn(d * k)])
        vp1 = np.concatenate([rs.randn(d), 0.4 * rs.randn(d) - 0.3, 0.5 * rs.randn(d * k)])
        with Recorder() as rec:
            x = fam.sample(vp, 40)
        assert rec[0][0] == 'randn' and rec[1][0] == 'randn'
        out[tag + '/z'] = rec[0][2]
        out[tag + '/eps'] = rec[1][2]
        out[tag + '/var_param'] = vp
        out[tag + '/var_param1'] = vp1
        out[tag + '/sample'] = x
        out[tag + '/log_density'] = np.asarray(fam.log_density(vp, x))
        out[tag + '/log_density_1d'] = np.asarray(fam.log_density(vp, x[0]))
        out[tag + '/entropy'] = np.asarray(fam.entropy(vp))
        out[tag + '/kl'] = np.asarray(fam.kl(vp, vp1))
        mean, cov = fam.mean_and_cov(vp)
        out[tag + '/mean'] = mean
        out[tag + '/cov'] = cov
        for p in (2, 4):
            out[tag + '/moment%d' % p] = np.asarray(fam.pth_moment(vp, p))
        out[tag + '/init_head'] = fam.init_param()[:2 * d]
    models = {}
    X, y, _ = logistic_problem(60, 4, seed=11)
    models['logistic_d4'] = (4, logistic_log_p(X, y, 10.0))
    mean, sd = target_params(5, seed=14)
    models['gauss_d5'] = (5, gauss_log_p(mean, sd))
    for mname, (d, logp) in models.items():
        for k in (0, 2):
            fam = LRGaussian(d, seed=1214, k=k)
            vp = np.concatenate([rs.randn(d), 0.4 * rs.randn(d) - 0.3, 0.5 * rs.randn(d * k)])
            objs = {'ekl': lambda: ExclusiveKL(fam, logp, 7), 'ekl_path': lambda: ExclusiveKL(fam, logp, 7, use_path_deriv=True),
                    'alpha2': lambda: AlphaDivergence(fam, logp, 7, 2.0)}
            for oname, mk in objs.items():
                tag = 'obj/%s/k%d/%s' % (mname, k, oname)
                np.random.seed(5039)
                with Recorder() as rec:
                    value, grad = mk()(vp)
                out[tag + '/z'] = rec[0][2]
                out[tag + '/eps'] = rec[1][2]
                out[tag + '/var_param'] = vp
                out[tag + '/value'] = np.asarray(float(value))
                out[tag + '/grad'] = np.asarray(grad, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, 'lr_gaussian.npz'), **out)
    print('lr_gaussian: %d arrays' % len(out))


def gen_optimizers():
    out = {}
    rs = np.random.RandomState(153)
    grads = rs.randn(6, 5) * np.array([1.0, 10.0, 0.1, 3.0, 1e-3])
    out['grads'] = grads
    for name, opt in (('rmsprop', RMSProp(0.01)), ('adam', Adam(0.01)),
                      ('rmsprop_b', RMSProp(0.01, beta=0.5, jitter=1e-6)),
                      ('adam_b', Adam(0.01, beta1=0.7, beta2=0.9, jitter=1e-6))):
        dirs = []
        for g in grads:
            dirs.append(np.array(opt.descent_direction(g.copy()), copy=True))
        out[name + '/dirs'] = np.array(dirs)

    # a short full optimize() run on a deterministic quadratic objective
    class Quad(object):
        def __call__(self, vp):
            return 0.5 * np.sum(vp ** 2), vp.copy()

        def update(self, vp, direction):
            return vp - direction
    for name, mk in (('rmsprop', lambda: RMSProp(0.1)), ('adam', lambda: Adam(0.1))):
        res = mk().optimize(25, Quad(), np.array([1.0, -2.0, 0.5]))
        out[name + '/opt_param'] = res['opt_param']
        out[name + '/value_history'] = res['value_history']
        out[name + '/param_history'] = res['variational_param_history']
    np.savez_compressed(os.path.join(OUT, 'optimizers.npz'), **out)
    print('optimizers: %d arrays' % len(out))


def gen_psis():
    out = {}
    for name in PSIS_CASES:
        lw = psis_case(name)
        res, k = ref_psis.psislw(lw.copy())
        n = lw.shape[0]
        out[name + '/khat'] = np.asarray(k)
        if lw.ndim == 1:
            # tail = entries above the cutoff, recomputed exactly as _psis.py:163-175 does
            x = lw - np.max(lw)
            M = int(np.ceil(min(0.2 * n, 3 * np.sqrt(n))))
            cut = max(np.sort(x)[-M - 1], np.log(np.finfo(float).tiny))
            out[name + '/tail_idx'] = np.flatnonzero(x > cut).astype(np.int64)
        stride = max(1, n // 4096)
        out[name + '/out_stride'] = np.asarray(stride)
        out[name + '/out_sub'] = res[::stride]
        out[name + '/out_sorted_sub'] = np.sort(res, axis=0)[::stride]
        out[name + '/out_max'] = np.max(res, axis=0)
        out[name + '/out_min'] = np.min(res, axis=0)
        out[name + '/out_mean'] = np.mean(res, axis=0)
        if lw.ndim == 1 and np.isfinite(k):
            d2, lnb = ref_diag.divergence_bound(res, return_log_norm_bound=True)
            out[name + '/d2'] = np.asarray(d2)
            out[name + '/elbo'] = np.asarray(lnb)
            for alpha in (1.5, 3.0):
                out[name + '/dalpha%.1f' % alpha] = np.asarray(
                    ref_diag.divergence_bound(res, alpha=alpha))
    # gpdfitnew / gpinv / sumlogs on their own
    rs = np.random.RandomState(1639)
    for n in (5, 37, 1000):
        x = np.sort(rs.pareto(2.5, size=n))
        k, sigma = ref_psis.gpdfitnew(x, sort=False)
        out['gpd_n%d/x' % n] = x
        out['gpd_n%d/k' % n] = np.asarray(k)
        out['gpd_n%d/sigma' % n] = np.asarray(sigma)
    p = (np.arange(0.5, 50) / 50)
    for k in (0.5, -0.3, 0.0):
        out['gpinv_k%.1f' % k] = ref_psis.gpinv(p, k, 1.7)
    v = rs.randn(1000) * 30
    out['sumlogs/x'] = v
    out['sumlogs/out'] = np.asarray(ref_psis.sumlogs(v))
    np.savez_compressed(os.path.join(OUT, 'psis.npz'), **out)
    print('psis: %d arrays' % len(out))


def gen_psisloo():
    rs = np.random.RandomState(4711)
    n, m = 4000, 5
    # Gaussian log-likelihood terms of different widths: raw weights exp(c z^2 / 2) with tail index k ~ c
    log_lik = -0.5 * np.linspace(0.1, 0.9, m)[None, :] * rs.randn(n, m) ** 2 - rs.gamma(2.0, 0.3, size=(1, m))
    loo, loos, ks = ref_psis.psisloo(log_lik.copy())
    np.savez_compressed(os.path.join(OUT, 'psisloo.npz'), log_lik=log_lik, loo=loo, loos=loos, ks=ks)
    print('psisloo: 4 arrays')


def gen_diagnostics():
    out = {}
    samples, lw = diag_problem()
    for alpha in (1.5, 2.0, 3.0):
        out['dalpha%.1f' % alpha] = np.asarray(ref_diag.divergence_bound(lw, alpha=alpha))
        out['dalpha%.1f_lnb0' % alpha] = np.asarray(
            ref_diag.divergence_bound(lw, alpha=alpha, log_norm_bound=0.0))
    wb = ref_diag.wasserstein_bounds(0.7, samples=samples)
    out['wb_samples'] = np.array([wb['W1'], wb['W2']])
    wb = ref_diag.wasserstein_bounds(0.7, samples=samples[:, 0])
    out['wb_samples_1d'] = np.array([wb['W1'], wb['W2']])
    wb = ref_diag.wasserstein_bounds(0.7, moment_bound_fn=lambda p: 3.0 * p)
    out['wb_fn'] = np.array([wb['W1'], wb['W2']])
    keys = ['W1', 'W2', 'mean_error', 'std_error', 'cov_error', 'd2', 'log_norm_bound']
    res = ref_diag.all_diagnostics(lw, samples=samples)
    out['all_samples'] = np.array([res[k] for k in keys])
    res = ref_diag.all_diagnostics(lw, moment_bound_fn=lambda p: 2.5 * p, q_var=1.7)
    out['all_fn_scalar'] = np.array([res[k] for k in keys])
    res = ref_diag.all_diagnostics(lw, samples=samples, q_var=np.cov(samples.T) * 1.1,
                                   p_var=0.9, log_norm_bound=-1.5)
    out['all_full'] = np.array([res[k] for k in keys])
    eb = ref_diag.error_bounds(W1=0.3, W2=0.5, q_var=2.0)
    out['error_bounds'] = np.array([eb['mean_error'], eb['std_error'], eb['cov_error']])

    # vi_diagnostics end to end (convenience.py:97-179) on MFGaussian vs Gaussian targets
    import contextlib
    import io
    from viabel import convenience
    from viabel.models import Model
    mean, sd = target_params(4, seed=21)
    for name, scale in (('matched', 1.05), ('narrow', 0.5), ('wide', 3.0)):
        fam = MFGaussian(4, seed=56)        # same seed each time -> identical eps
        vp = np.concatenate([mean, np.log(sd)])
        model = Model(gauss_log_p(mean + 0.02, sd * scale))
        with Recorder() as rec, contextlib.redirect_stdout(io.StringIO()):
            res = convenience.vi_diagnostics(vp, model=model, approx=fam, n_samples=20000)
        out['vi_eps'] = rec[0][2]
        out['vi_%s/var_param' % name] = vp
        out['vi_%s/target_mean' % name] = mean + 0.02
        out['vi_%s/target_sd' % name] = sd * scale
        out['vi_%s/khat' % name] = np.asarray(res['khat'])
        out['vi_%s/slw_sub' % name] = res['smoothed_log_weights'][::20]
        for k in keys:
            if k in res:
                out['vi_%s/%s' % (name, k)] = np.asarray(res[k])
    np.savez_compressed(os.path.join(OUT, 'diagnostics.npz'), **out)
    print('diagnostics: %d arrays' % len(out))


def mc_chains():
    """Synthetic optimisation traces for the convergence statistics FASO / RAABBVI compute on the host
    (optimization.py:550-605): an AR(1) iterate window with drift, [n_iters, n_params]."""
    rs = np.random.RandomState(4242)
    n, P = 600, 5
    x = np.zeros((n, P))
    phi = np.array([0.0, 0.5, 0.9, 0.97, -0.4])
    e = rs.randn(n, P)
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    x += np.linspace(0.0, 1.0, n)[:, None] * np.array([0.0, 0.0, 0.0, 2.0, 0.0])      # one drifting coordinate
    return x


def gen_mc_diagnostics():
    from viabel import _mc_diagnostics as ref_mc
    out = {}
    x = mc_chains()
    out['acov'] = ref_mc.autocov(x[:, :3].T)
    out['ess_1chain'] = np.array([ref_mc.ess(x[:, j][None, :]) for j in range(x.shape[1])])
    out['ess_4chains'] = np.array([ref_mc.ess(x[:, j].reshape(4, -1)) for j in range(x.shape[1])])
    eff, mcse = ref_mc.MCSE(x)
    out['mcse_ess'] = np.asarray(eff, dtype=np.float64)
    out['mcse'] = np.asarray(mcse, dtype=np.float64)
    out['rhat'] = np.asarray(ref_mc.compute_R_hat(x))
    out['rhat_warm_odd'] = np.asarray(ref_mc.compute_R_hat(x, warmup=51))
    windows = np.array([100, 200, 300, 450])
    for tag, cols in (('stationary', [0, 1, 4]), ('drifting', [0, 3])):
        ok, best_W = ref_mc.R_hat_convergence_check(x[:, cols], windows)
        out['check_%s' % tag] = np.array([float(ok), float(best_W)])
    np.savez_compressed(os.path.join(OUT, 'mc_diagnostics.npz'), **out)
    print('mc_diagnostics: %d arrays' % len(out))


def dis_inputs():
    """Fixed (samples, log p, log q, temper prior) triples for the ESS bisection of DISInclusiveKL."""
    rs = np.random.RandomState(515)
    S, d = 400, 3
    x = rs.randn(S, d) * 1.5
    cases = {}
    log_q = -0.5 * np.sum((x / 1.5) ** 2, axis=1) - d * np.log(1.5 * np.sqrt(2 * np.pi))
    # target much narrower than q: the ESS target is reached strictly inside (0, 1)
    cases['interior'] = (x, -0.5 * np.sum((x / 0.4) ** 2, axis=1), log_q, 120)
    # target close to q: even eps = 0 (no tempering) satisfies the ESS target -> eps snaps to 0
    cases['eps0'] = (x, -0.5 * np.sum((x / 1.4) ** 2, axis=1), log_q, 60)
    # unreachable ESS target: eps stays at its upper end point 1
    cases['eps1'] = (x, -0.5 * np.sum((x / 0.2) ** 2, axis=1), log_q, 399)
    return cases


def gen_dis():
    from viabel.models import Model
    from viabel.objectives import DISInclusiveKL
    out = {}
    for name, (x, log_p, log_q, target) in dis_inputs().items():
        d = x.shape[1]
        prior = MFGaussian(d)
        obj = DISInclusiveKL(MFGaussian(d), Model(lambda z: -0.5 * anp.sum(z ** 2, axis=1)), x.shape[0], target,
                             prior, np.zeros(2 * d))
        eps, ess, w = obj._get_eps_and_weights(1, x, log_p, log_q)
        out['%s/eps' % name] = np.asarray(float(eps))
        out['%s/ess' % name] = np.asarray(float(ess))
        out['%s/w' % name] = np.asarray(w, dtype=np.float64)
        out['%s/log_prior' % name] = np.asarray(prior.log_density(np.zeros(2 * d), x), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, 'dis.npz'), **out)
    print('dis: %d arrays' % len(out))


def gen_dis_objective():
    """DISInclusiveKL end to end (objectives.py:283-416): objective value and gradient of consecutive calls, with and
    without resampling, through the reference's own value_and_grad.  The resampling indices come from numpy's GLOBAL
    generator (:408-409), which the product also uses: seeding it identically reproduces them."""
    from viabel.objectives import DISInclusiveKL
    out = {}
    mean, sd = target_params(4, seed=14)
    logp = gauss_log_p(mean, sd)
    rs = np.random.RandomState(5039)
    for kind, df in (('mfg', None), ('mft', 8)):
        for resample, batches in ((False, 1), (True, 1), (True, 3)):
            tag = 'dis_obj/%s/%s%d' % (kind, 'res' if resample else 'nores', batches)
            fam = make_family(kind, 4, df, seed=1214)
            prior = MFGaussian(4)
            obj = DISInclusiveKL(fam, logp, 60, 20, prior, np.concatenate([np.zeros(4), np.ones(4)]),
                                 use_resampling=resample, num_resampling_batches=batches)
            vp = np.concatenate([mean + 0.3 * rs.randn(4), np.log(sd) + 0.2 * rs.randn(4)])
            np.random.seed(846)
            n_calls = 4
            for c in range(n_calls):
                with Recorder() as rec:
                    value, grad = obj(vp)
                if len(rec):
                    out['%s/call%d/base' % (tag, c)] = rec[0][2]
                out['%s/call%d/var_param' % (tag, c)] = vp
                out['%s/call%d/value' % (tag, c)] = np.asarray(float(value))
                out['%s/call%d/grad' % (tag, c)] = np.asarray(grad, dtype=np.float64)
                out['%s/call%d/eps' % (tag, c)] = np.asarray(float(obj._eps))
                vp = vp - 0.05 * np.asarray(grad)
    np.savez_compressed(os.path.join(OUT, 'dis_objective.npz'), **out)
    print('dis_objective: %d arrays' % len(out))


def nvp_masks(dim, n_pairs):
    """Alternating half masks of the reference's own NVP test (tests/test_approximations.py:131-139)."""
    half, halfplus = dim // 2, dim - dim // 2
    m1 = np.hstack([[0] * half, [1] * halfplus])
    m2 = np.hstack([[1] * half, [0] * halfplus])
    return np.array(list(np.vstack([m1, m2])) * n_pairs)


def gen_flows():
    """NeuralNet.forward (approximations.py:414-429) and NVPFlow (:452-550): g / f / log_density / sample with the
    prior's recorded draws, then ExclusiveKL(use_path_deriv=True) and AlphaDivergence on the flow (the only two
    objectives the reference can evaluate on a family without entropy: its plain ExclusiveKL branch calls
    approx.log_density(samples) without var_param, objectives.py:163)."""
    from viabel.approximations import NeuralNet, NVPFlow
    out = {}
    rs = np.random.RandomState(1639)
    for dim, hidden in ((1, 4), (3, 10), (6, 8)):
        tag = 'nn_d%d' % dim
        nn = NeuralNet([[dim, hidden], [hidden, hidden], [hidden, dim]])
        flat = rs.randn(nn.var_param_dim) / 3
        x = rs.randn(9, dim)
        y, ldj = nn.forward(nn._pattern.fold(flat), x)
        out[tag + '/flat'] = flat
        out[tag + '/x'] = x
        out[tag + '/y'] = np.asarray(y)
        out[tag + '/log_det_J'] = np.asarray(ldj)
    for dim, hidden, pairs, prior_kind in ((1, 5, 2, 'mfg'), (3, 10, 3, 'mfg'), (6, 7, 2, 'mft')):
        tag = 'nvp_d%d' % dim
        layers = [[dim, hidden], [hidden, dim]]
        mask = nvp_masks(dim, pairs)
        prior = MFGaussian(dim) if prior_kind == 'mfg' else MFStudentT(dim, 7)
        prior_param = np.concatenate([0.2 * rs.randn(dim), 0.1 * rs.randn(dim)])
        fam = NVPFlow(layers, layers, mask, prior, prior_param, dim)
        vp = rs.randn(fam.var_param_dim) / 4
        with Recorder() as rec:
            x = fam.sample(vp, 11, seed=341)
        z0 = prior_param[:dim] + np.exp(prior_param[dim:]) * rec[0][2]
        out[tag + '/prior_param'] = prior_param
        out[tag + '/mask'] = mask.astype(np.float64)
        out[tag + '/var_param'] = vp
        out[tag + '/base'] = rec[0][2]
        out[tag + '/z0'] = z0
        out[tag + '/sample'] = np.asarray(x)
        zb, ld = fam.f(vp, x)
        out[tag + '/f_z'] = np.asarray(zb)
        out[tag + '/f_logdet'] = np.asarray(ld)
        out[tag + '/log_density'] = np.asarray(fam.log_density(vp, x))
        mean, sd = target_params(dim, seed=14 + dim)
        logp = gauss_log_p(mean, sd)
        objs = {'ekl_path': lambda: ExclusiveKL(fam, logp, 8, use_path_deriv=True),
                'alpha2': lambda: AlphaDivergence(fam, logp, 8, 2.0)}
        for oname, mk in objs.items():
            otag = '%s/obj/%s' % (tag, oname)
            np.random.seed(5039)
            with Recorder() as rec:
                value, grad = mk()(vp)
            out[otag + '/base'] = rec[0][2]
            out[otag + '/value'] = np.asarray(float(value))
            out[otag + '/grad'] = np.asarray(grad, dtype=np.float64)
        out[tag + '/target_mean'] = mean
        out[tag + '/target_sd'] = sd
    np.savez_compressed(os.path.join(OUT, 'flows.npz'), **out)
    print('flows: %d arrays' % len(out))


if __name__ == '__main__':
    print('reference:', os.path.dirname(viabel.__file__))
    gen_families()
    gen_objectives()
    gen_objectives_cv()
    gen_lr_gaussian()
    gen_optimizers()
    gen_psis()
    gen_psisloo()
    gen_diagnostics()
    gen_mc_diagnostics()
    gen_dis()
    gen_dis_objective()
    gen_flows()
