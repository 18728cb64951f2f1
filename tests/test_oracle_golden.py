"""CPU: the numpy oracle (oracle/viabel_oracle.py) against the golden vectors produced
by the unmodified reference (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import relerr
from _problems import (PSIS_CASES, diag_problem, hier_problem, logistic_problem, psis_case,
                       target_params)
from oracle import viabel_oracle as vo

TOL = 1e-10


def _split(tag):
    kind, df, d = tag.split('_')
    df = None if df == 'dfNone' else float(df[2:])
    return kind, df, int(d[1:])


FAMILY_TAGS = ['%s_df%s_d%d' % (k, df, d)
               for k, df in (('mfg', None), ('mft', 20), ('mft', 5.5), ('mvt', 100), ('mvt', 7))
               for d in (1, 3, 8)]


@pytest.mark.parametrize('tag', FAMILY_TAGS)
def test_family(golden, tag):
    g = golden('families')
    kind, df, d = _split(tag)
    vp, vp1 = g[tag + '/var_param'], g[tag + '/var_param1']
    if kind == 'mvt':
        x = vo.mvt_sample(vp, g[tag + '/chi2'], g[tag + '/z'], df)
        logq = vo.mvt_log_density(vp, x, df)
        logq1 = vo.mvt_log_density(vp, x[0], df)
        ent = vo.mvt_entropy(vp, d)
        assert relerr(vo.mvt_entropy_stable(vp, d), g[tag + '/entropy']) < TOL
        mean, cov = vo.mvt_mean_and_cov(vp, d, df)
        mom = lambda p: vo.mvt_pth_moment(vp, d, df, p)
        init = vo.mvt_init_param(d)
    elif kind == 'mft':
        x = vo.mft_sample(vp, g[tag + '/base'])
        logq = vo.mft_log_density(vp, x, df)
        logq1 = vo.mft_log_density(vp, x[0], df)
        ent = vo.mft_entropy(vp, d)
        mean, cov = vo.mft_mean_and_cov(vp, d, df)
        mom = lambda p: vo.mft_pth_moment(vp, d, df, p)
        init = vo.mfg_init_param(d)
    else:
        x = vo.mfg_sample(vp, g[tag + '/base'])
        logq = vo.mfg_log_density(vp, x)
        logq1 = vo.mfg_log_density(vp, x[0])
        ent = vo.mfg_entropy(vp, d)
        mean, cov = vo.mfg_mean_and_cov(vp, d)
        mom = lambda p: vo.mfg_pth_moment(vp, d, p)
        init = vo.mfg_init_param(d)
        assert relerr(vo.mfg_kl(vp, vp1, d), g[tag + '/kl']) < TOL
    assert relerr(init, g[tag + '/init_param']) < TOL
    assert relerr(x, g[tag + '/sample']) < TOL
    assert relerr(logq, g[tag + '/log_density']) < TOL
    assert relerr(logq1, g[tag + '/log_density_1d']) < TOL
    assert relerr(ent, g[tag + '/entropy']) < TOL
    assert relerr(mean, g[tag + '/mean']) < TOL
    assert relerr(cov, g[tag + '/cov']) < TOL
    for p in (2, 4):
        key = tag + '/moment%d' % p
        if key in g:
            assert relerr(mom(p), g[key]) < TOL
        else:
            with pytest.raises(ValueError):
                mom(p)


def _models():
    m = {}
    X, y, _ = logistic_problem(60, 4, seed=11)
    m['logistic_d4'] = lambda th, X=X, y=y: vo.logistic_logp_grad(th, X, y, 10.0)
    X, y, _ = logistic_problem(1000, 10, seed=12)
    m['logistic_d10'] = lambda th, X=X, y=y: vo.logistic_logp_grad(th, X, y, 10.0)
    X, y, _ = logistic_problem(200, 6, seed=13)
    m['probit_d6'] = lambda th, X=X, y=y: vo.probit_logp_grad(th, X, y, 10.0)
    mean, sd = target_params(5, seed=14)
    m['gauss_d5'] = lambda th: vo.gauss_target_logp_grad(th, mean, sd)
    m['student_d5'] = lambda th: vo.student_target_logp_grad(th, mean, sd, 10.0)
    hp = hier_problem(G=3, p=2, n_per=7, seed=15)
    m['hier_G3p2'] = lambda th: vo.hier_linear_logp_grad(th, hp['X'], hp['y'], hp['group'], 3, 2)
    return m


def _objective_tags(g):
    return sorted({k.rsplit('/', 1)[0] for k in g if k.endswith('/value')})


def test_objectives_all(golden):
    g = golden('objectives')
    models = _models()
    tags = _objective_tags(g)
    assert len(tags) > 80
    worst = 0.0
    for tag in tags:
        mname, fam, point, oname = tag.split('/')
        kind, df = fam.split('_df')
        df = None if df == 'None' else float(df)
        vp = g[tag + '/var_param']
        model = models[mname]
        if kind == 'mvt':
            chi2, z = g[tag + '/chi2'], g[tag + '/z']
            if oname == 'ekl':
                v, gr, _ = vo.exclusive_kl_mvt(vp, chi2, z, model, df)
            else:
                v, gr, _ = vo.alpha_divergence_mvt(vp, chi2, z, model, df, float(oname[5:]))
        else:
            family = 'gaussian' if kind == 'mfg' else 'student'
            base = g[tag + '/base']
            if oname.startswith('ekl'):
                v, gr, _ = vo.exclusive_kl_meanfield(vp, base, model, family, df,
                                                     path_deriv=oname.endswith('path'))
            else:
                v, gr, _ = vo.alpha_divergence_meanfield(vp, base, model, float(oname[5:]),
                                                         family, df)
        ev, eg = relerr(v, g[tag + '/value']), relerr(gr, g[tag + '/grad'])
        worst = max(worst, ev, eg)
        assert ev < TOL, (tag, 'value', ev)
        assert eg < TOL, (tag, 'grad', eg)
    print('worst objective rel err', worst)


def test_optimizer_directions(golden):
    g = golden('optimizers')
    grads = g['grads']
    for name, fn, kw in (('rmsprop', vo.rmsprop_direction, {}),
                         ('adam', vo.adam_direction, {}),
                         ('rmsprop_b', vo.rmsprop_direction, dict(beta=0.5, jitter=1e-6)),
                         ('adam_b', vo.adam_direction, dict(beta1=0.7, beta2=0.9, jitter=1e-6))):
        st = {}
        for i, gr in enumerate(grads):
            d = fn(st, gr, **kw)
            assert relerr(d, g[name + '/dirs'][i]) < 1e-13, (name, i)


@pytest.mark.parametrize('name', PSIS_CASES)
def test_psis(golden, name):
    g = golden('psis')
    lw = psis_case(name)
    with np.errstate(all='ignore'):
        out, k = vo.psislw(lw)
    kref = g[name + '/khat']
    assert np.array_equal(np.isinf(k), np.isinf(kref))
    fin = np.isfinite(kref)
    assert relerr(np.asarray(k)[fin], kref[fin]) < 1e-12
    stride = int(g[name + '/out_stride'])
    if name != 'ties_3e4':
        # with tied tail values the reference's unstable argsort (_psis.py:183) hands the
        # smoothed quantiles to tied entries in an unspecified order; only the multiset is pinned
        np.testing.assert_allclose(out[::stride], g[name + '/out_sub'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.sort(out, axis=0)[::stride], g[name + '/out_sorted_sub'],
                               rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.max(out, axis=0), g[name + '/out_max'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.min(out, axis=0), g[name + '/out_min'], rtol=1e-12, atol=1e-12)
    if lw.ndim == 1:
        _, _, tail, _ = vo.psislw_1d(lw, return_tail=True)
        assert np.array_equal(tail, g[name + '/tail_idx'])        # bit-exact index set
        # the argsort-structured variant (the CPU baseline bench.py times) gives the same numbers
        with np.errstate(all='ignore'):
            out2, k2 = vo.psislw_argsort_1d(lw)
        assert (np.isinf(k2) and np.isinf(k)) or relerr(k2, k) < 1e-13
        np.testing.assert_allclose(np.sort(out2), np.sort(out), rtol=1e-13, atol=1e-13)
    if name + '/d2' in g:
        d2, elbo, _ = vo.divergence_bound(out)
        assert relerr(d2, g[name + '/d2']) < 1e-10
        assert relerr(elbo, g[name + '/elbo']) < 1e-12
        for alpha in (1.5, 3.0):
            assert relerr(vo.divergence_bound(out, alpha)[0], g[name + '/dalpha%.1f' % alpha]) < 1e-10


def test_gpd_helpers(golden):
    g = golden('psis')
    for n in (5, 37, 1000):
        with np.errstate(all='ignore'):
            k, sigma = vo.gpdfit(g['gpd_n%d/x' % n])
        assert relerr(k, g['gpd_n%d/k' % n]) < 1e-12
        assert relerr(sigma, g['gpd_n%d/sigma' % n]) < 1e-12
    p = np.arange(0.5, 50) / 50
    for k in (0.5, -0.3, 0.0):
        assert relerr(vo.gpinv(p, k, 1.7), g['gpinv_k%.1f' % k]) < 1e-13
    assert relerr(vo.sumlogs(g['sumlogs/x']), g['sumlogs/out']) < 1e-14
    with pytest.raises(ValueError):
        vo.psislw(np.zeros(1))
    with pytest.raises(ValueError):
        vo.psislw(np.zeros((2, 2, 2)))


def test_psisloo(golden):
    g = golden('psisloo')
    with np.errstate(all='ignore'):
        loo, loos, ks = vo.psisloo(g['log_lik'])
    assert relerr(loo, g['loo']) < TOL and relerr(loos, g['loos']) < TOL and relerr(ks, g['ks']) < TOL


def test_diagnostics(golden):
    g = golden('diagnostics')
    samples, lw = diag_problem()
    keys = ['W1', 'W2', 'mean_error', 'std_error', 'cov_error', 'd2', 'log_norm_bound']
    for alpha in (1.5, 2.0, 3.0):
        assert relerr(vo.divergence_bound(lw, alpha)[0], g['dalpha%.1f' % alpha]) < TOL
        assert relerr(vo.divergence_bound(lw, alpha, 0.0)[0], g['dalpha%.1f_lnb0' % alpha]) < TOL
    wb = vo.wasserstein_bounds(0.7, samples=samples)
    assert relerr([wb['W1'], wb['W2']], g['wb_samples']) < TOL
    wb = vo.wasserstein_bounds(0.7, samples=samples[:, 0])
    assert relerr([wb['W1'], wb['W2']], g['wb_samples_1d']) < TOL
    wb = vo.wasserstein_bounds(0.7, moment_bound_fn=lambda p: 3.0 * p)
    assert relerr([wb['W1'], wb['W2']], g['wb_fn']) < TOL
    res = vo.all_diagnostics(lw, samples=samples)
    assert relerr([res[k] for k in keys], g['all_samples']) < TOL
    res = vo.all_diagnostics(lw, moment_bound_fn=lambda p: 2.5 * p, q_var=1.7)
    assert relerr([res[k] for k in keys], g['all_fn_scalar']) < TOL
    res = vo.all_diagnostics(lw, samples=samples, q_var=np.cov(samples.T) * 1.1, p_var=0.9,
                             log_norm_bound=-1.5)
    assert relerr([res[k] for k in keys], g['all_full']) < TOL
    eb = vo.error_bounds(W1=0.3, W2=0.5, q_var=2.0)
    assert relerr([eb['mean_error'], eb['std_error'], eb['cov_error']], g['error_bounds']) < TOL
    with pytest.raises(ValueError):
        vo.divergence_bound(lw, alpha=1.0)
    with pytest.raises(ValueError):
        vo.wasserstein_bounds(0.5)


def test_vi_diagnostics_pipeline(golden):
    """convenience.py:136-179 restated with the oracle pieces."""
    g = golden('diagnostics')
    eps = g['vi_eps']
    for name in ('matched', 'narrow', 'wide'):
        vp = g['vi_%s/var_param' % name]
        x = vo.mfg_sample(vp, eps)
        lp, _ = vo.gauss_target_logp_grad(x, g['vi_%s/target_mean' % name], g['vi_%s/target_sd' % name])
        lw = lp - vo.mfg_log_density(vp, x)
        with np.errstate(all='ignore'):
            slw, k = vo.psislw(lw)
        assert relerr(k, g['vi_%s/khat' % name]) < 1e-9
        np.testing.assert_allclose(slw[::20], g['vi_%s/slw_sub' % name], rtol=1e-9, atol=1e-9)
        if k > 0.7:
            assert 'vi_%s/d2' % name not in g
            continue
        res = vo.all_diagnostics(slw, samples=x.T, moment_bound_fn=lambda p: vo.mfg_pth_moment(vp, 4, p),
                                 q_var=vo.mfg_mean_and_cov(vp, 4)[1])
        for key in ('W1', 'W2', 'mean_error', 'std_error', 'cov_error', 'd2', 'log_norm_bound'):
            assert relerr(res[key], g['vi_%s/%s' % (name, key)]) < 1e-9, (name, key)


@pytest.mark.parametrize('name,cuts', [('t5_t7_1e5', [0, 50000, 100000]), ('t5_t7_1e5', [0, 700, 61234, 100000]),
                                       ('small_100', [0, 10, 22, 100]), ('ties_3e4', [0, 15000, 30000]),
                                       ('underflow_5000', [0, 2500, 5000]), ('all_equal_50', [0, 20, 50])])
def test_sharded_psis_rule_equals_single(name, cuts):
    """The record rule of the draw-sharded PSIS (each rank ships its top M+1; csrc/psis.cu) gives
    the same smoothed weights and k-hat as the one-column algorithm pinned against _psis.py."""
    lw = psis_case(name)
    ref, kref = vo.psislw_1d(lw)
    outs, k = vo.psislw_sharded([lw[cuts[i]:cuts[i + 1]] for i in range(len(cuts) - 1)])
    assert (np.isinf(k) and np.isinf(kref)) or k == kref
    np.testing.assert_allclose(np.concatenate(outs), ref, rtol=0, atol=1e-11)


def _mc_chains():
    """Same traces as oracle/make_golden.py::mc_chains (rebuilt by seed)."""
    rs = np.random.RandomState(4242)
    n, P = 600, 5
    x = np.zeros((n, P))
    phi = np.array([0.0, 0.5, 0.9, 0.97, -0.4])
    e = rs.randn(n, P)
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    x += np.linspace(0.0, 1.0, n)[:, None] * np.array([0.0, 0.0, 0.0, 2.0, 0.0])
    return x


def test_mc_diagnostics_golden(golden):
    """Host-side convergence statistics of FASO / RAABBVI (viabel_b200/_mc_diagnostics.py, a numpy mirror of
    viabel/_mc_diagnostics.py:7-184) against the unmodified reference: autocovariance, ESS, MCSE,
    split R-hat and the window check."""
    from viabel_b200 import _mc_diagnostics as mc
    g = golden('mc_diagnostics')
    x = _mc_chains()
    np.testing.assert_allclose(mc.autocov(x[:, :3].T), g['acov'], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose([mc.ess(x[:, j][None, :]) for j in range(5)], g['ess_1chain'], rtol=1e-10)
    np.testing.assert_allclose([mc.ess(x[:, j].reshape(4, -1)) for j in range(5)], g['ess_4chains'], rtol=1e-10)
    eff, mcse = mc.MCSE(x)
    np.testing.assert_allclose(np.asarray(eff, dtype=np.float64), g['mcse_ess'], rtol=1e-10)
    np.testing.assert_allclose(np.asarray(mcse, dtype=np.float64), g['mcse'], rtol=1e-10)
    np.testing.assert_allclose(mc.compute_R_hat(x), g['rhat'], rtol=1e-12)
    np.testing.assert_allclose(mc.compute_R_hat(x, warmup=51), g['rhat_warm_odd'], rtol=1e-12)
    windows = np.array([100, 200, 300, 450])
    for tag, cols in (('stationary', [0, 1, 4]), ('drifting', [0, 3])):
        ok, best = mc.R_hat_convergence_check(x[:, cols], windows)
        assert [float(ok), float(best)] == g['check_%s' % tag].tolist()


def test_sharded_psis_rule_random_partitions_and_ties():
    """Property test of the record rule on tie-heavy data (values on a coarse grid, so the cutoff value is shared
    by many draws on several ranks) and random, very uneven partitions -- including ranks with fewer than M+1
    draws."""
    rs = np.random.RandomState(77)
    for trial in range(12):
        n = int(rs.choice([300, 2000, 20000]))
        x = np.round(rs.standard_t(3, n) * 2.0, 1 if trial % 2 else 2) - 5.0
        R = int(rs.choice([2, 3, 5, 8]))
        cuts = np.sort(rs.choice(np.arange(2, n - 2), R - 1, replace=False))
        parts = np.split(x, cuts)
        if min(len(p) for p in parts) < 2:
            continue
        ref, kref = vo.psislw_1d(x)
        outs, k = vo.psislw_sharded(parts)
        assert (np.isinf(k) and np.isinf(kref)) or abs(k - kref) <= 1e-12 * abs(kref), (trial, k, kref)
        np.testing.assert_allclose(np.concatenate(outs), ref, rtol=0, atol=1e-11)


CV_MODELS = ['logistic_d4', 'logistic_d10', 'probit_d6', 'gauss_d5', 'student_d5']


def cv_model(name):
    """(model(theta) -> (f, G), hessian(m) -> H) of the control-variate golden cases (oracle/make_golden.py)."""
    if name.startswith('logistic'):
        N, d, seed = (60, 4, 11) if name == 'logistic_d4' else (1000, 10, 12)
        X, y, _ = logistic_problem(N, d, seed=seed)
        return (lambda th: vo.logistic_logp_grad(th, X, y, 10.0)), (lambda m: vo.logistic_hessian(m, X, y, 10.0))
    if name == 'probit_d6':
        X, y, _ = logistic_problem(200, 6, seed=13)
        return (lambda th: vo.probit_logp_grad(th, X, y, 10.0)), (lambda m: vo.probit_hessian(m, X, y, 10.0))
    mean, sd = target_params(5, seed=14)
    if name == 'gauss_d5':
        return (lambda th: vo.gauss_target_logp_grad(th, mean, sd)), (lambda m: vo.gauss_target_hessian(m, mean, sd))
    return (lambda th: vo.student_target_logp_grad(th, mean, sd, 10.0)), \
        (lambda m: vo.student_target_hessian(m, mean, sd, 10.0))


@pytest.mark.parametrize('mname', CV_MODELS)
def test_control_variate_objectives(golden, mname):
    """ExclusiveKL with hessian_approx_method (objectives.py:170-273): all four estimators x {plain, path-derivative}
    x {MFGaussian, MFStudentT} against the unmodified reference run through the autograd stand-in."""
    g = golden('objectives_cv')
    model, hess = cv_model(mname)
    n = 0
    for key in [k for k in g if k.startswith(mname + '/') and k.endswith('/value')]:
        tag = key[:-len('/value')]
        _, fam, method, mode = tag.split('/')
        family, df = ('gaussian', None) if fam.startswith('mfg') else ('student', 8.0)
        v, gr = vo.exclusive_kl_cv_meanfield(g[tag + '/var_param'], g[tag + '/base'], model, hess, method, family, df,
                                             mode == 'path')
        assert relerr(v, g[tag + '/value']) < 1e-12, tag
        assert relerr(gr, g[tag + '/grad']) < 1e-10, tag
        n += 1
    assert n == 16


def test_lr_gaussian(golden):
    """LRGaussian family functions and objectives (approximations.py:610-731) against the unmodified reference."""
    g = golden('lr_gaussian')
    for d, k in ((3, 0), (3, 1), (8, 3), (6, 6)):
        t = 'lr_d%d_k%d' % (d, k)
        vp, vp1 = g[t + '/var_param'], g[t + '/var_param1']
        x = vo.lr_sample(vp, g[t + '/z'], g[t + '/eps'])
        assert relerr(x, g[t + '/sample']) < 1e-13
        assert relerr(vo.lr_log_density(vp, x, k), g[t + '/log_density']) < 1e-11
        assert relerr(vo.lr_log_density(vp, x[0], k), g[t + '/log_density_1d']) < 1e-11
        assert relerr(vo.lr_entropy(vp, d, k), g[t + '/entropy']) < 1e-12
        assert relerr(vo.lr_kl(vp, vp1, d, k), g[t + '/kl']) < 1e-10
        mean, cov = vo.lr_mean_and_cov(vp, d, k)
        assert relerr(mean, g[t + '/mean']) < 1e-14 and relerr(cov, g[t + '/cov']) < 1e-13
        for p in (2, 4):
            assert relerr(vo.lr_pth_moment(vp, d, k, p), g[t + '/moment%d' % p]) < 1e-12
    models = {'logistic_d4': cv_model('logistic_d4')[0], 'gauss_d5': cv_model('gauss_d5')[0]}
    n = 0
    for key in [q for q in g if q.startswith('obj/') and q.endswith('/value')]:
        t = key[:-len('/value')]
        _, mname, kk, oname = t.split('/')
        kind = 'alpha' if oname == 'alpha2' else oname
        v, gr = vo.lr_objective(g[t + '/var_param'], g[t + '/z'], g[t + '/eps'], models[mname], kind, 2.0)
        assert relerr(v, g[t + '/value']) < 1e-11, t
        assert relerr(gr, g[t + '/grad']) < 1e-9, t
        n += 1
    assert n == 12


FLOW_NN = ((1, 4), (3, 10), (6, 8))
FLOW_NVP = ((1, 5, None), (3, 10, None), (6, 7, 7))


def test_flow_forward_functions(golden):
    """NeuralNet.forward and NVPFlow g / f / log_density (approximations.py:414-429, :493-535) against the reference."""
    g = golden('flows')
    for dim, hidden in FLOW_NN:
        t = 'nn_d%d' % dim
        y, ld = vo.nn_forward(g[t + '/flat'], [(dim, hidden), (hidden, hidden), (hidden, dim)], g[t + '/x'])
        assert relerr(y, g[t + '/y']) < TOL and relerr(ld, g[t + '/log_det_J']) < TOL
    for dim, hidden, df in FLOW_NVP:
        t = 'nvp_d%d' % dim
        sh = [(dim, hidden), (hidden, dim)]
        vp, mask = g[t + '/var_param'], g[t + '/mask']
        assert relerr(vo.nvp_g(vp, sh, sh, mask, g[t + '/z0']), g[t + '/sample']) < TOL
        z, ld = vo.nvp_f(vp, sh, sh, mask, g[t + '/sample'])
        assert relerr(z, g[t + '/f_z']) < TOL and relerr(ld, g[t + '/f_logdet']) < TOL
        assert relerr(z, g[t + '/z0']) < 1e-12                        # f inverts g
        lq = vo.nvp_log_density(vp, sh, sh, mask, g[t + '/sample'], g[t + '/prior_param'], df)
        assert relerr(lq, g[t + '/log_density']) < TOL
