"""GPU parity: the CUDA path (through the C ABI / viabel_b200 API) against the golden vectors
of the unmodified reference and against the numpy oracle on seeded inputs.
FP64 path tolerance: 1e-10 relative (norm-wise for vectors) -- BASELINE.json north_star."""
import math

import numpy as np
import pytest
import torch

from conftest import relerr
from _problems import hier_problem, logistic_problem, target_params

pytestmark = pytest.mark.gpu

TOL64 = 1e-10


@pytest.fixture(scope='module')
def vb():
    import viabel_b200
    return viabel_b200


@pytest.fixture(scope='module')
def vo():
    from oracle import viabel_oracle
    return viabel_oracle


def test_library_loaded(vb):
    assert torch.cuda.is_available()
    assert vb._lib.lib.vb_device_sm_count() > 0
    assert vb._lib.lib.vb_version() >= 100


def test_philox_normal(vb):
    fam = vb.MFGaussian(1000, seed=7)
    z = fam.base_draws(2000).cpu().numpy().ravel()
    assert abs(z.mean()) < 4 / np.sqrt(z.size)
    assert abs(z.var() - 1) < 0.01
    assert abs(np.mean(z ** 4) - 3) < 0.05
    # counter-based: a slice of the stream can be regenerated from its offset (even and odd)
    lib, ptr = vb._lib.lib, vb._lib.ptr
    full = torch.empty(1001, dtype=torch.float64, device='cuda')
    lib.vb_philox_normal_f64(ptr(full), 1001, 99, 0, 0, vb._lib.stream())
    for off, n in ((10, 100), (7, 64), (501, 500)):
        part = torch.empty(n, dtype=torch.float64, device='cuda')
        lib.vb_philox_normal_f64(ptr(part), n, 99, off, 0, vb._lib.stream())
        assert torch.equal(part, full[off:off + n])
    # a different seed gives a different stream; same seed is reproducible
    a = vb.MFGaussian(8, seed=3).base_draws(4)
    b = vb.MFGaussian(8, seed=3).base_draws(4)
    c = vb.MFGaussian(8, seed=4).base_draws(4)
    assert torch.equal(a, b) and not torch.equal(a, c)
    # bf16-quantised draws are exactly representable in bfloat16
    fam.quantize_draws = 1
    q = fam.base_draws(100)
    assert torch.equal(q, q.to(torch.bfloat16).to(torch.float64))


def test_philox_student_and_chisquare(vb):
    lib, ptr = vb._lib.lib, vb._lib.ptr
    n = 400000
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    for df in (0.7, 3.0, 20.0):
        vb._lib.check(lib.vb_philox_chisquare_f64(ptr(out), n, df, 5, 0, vb._lib.stream()))
        x = out.cpu().numpy()
        assert np.all(x > 0)
        assert abs(x.mean() - df) < 6 * np.sqrt(2 * df / n)
        assert abs(x.var() - 2 * df) < 0.05 * 2 * df
    df = 9.0
    vb._lib.check(lib.vb_philox_student_t_f64(ptr(out), n, df, 5, 0, 0, vb._lib.stream()))
    t = out.cpu().numpy()
    assert abs(t.mean()) < 5 * np.sqrt(df / (df - 2) / n)
    assert abs(t.var() - df / (df - 2)) < 0.03


FAMILY_TAGS = ['%s_df%s_d%d' % (k, df, d)
               for k, df in (('mfg', None), ('mft', 20), ('mft', 5.5), ('mvt', 100), ('mvt', 7)) for d in (1, 3, 8)]


@pytest.mark.parametrize('tag', FAMILY_TAGS)
def test_family_golden(vb, golden, tag):
    g = golden('families')
    kind, df, d = tag.split('_')
    d = int(d[1:])
    if kind == 'mvt':
        fam = vb.MultivariateT(d, float(df[2:]))
        base = (g[tag + '/chi2'], g[tag + '/z'])
    else:
        fam = vb.MFGaussian(d) if kind == 'mfg' else vb.MFStudentT(d, float(df[2:]))
        base = g[tag + '/base']
    vp, vp1 = g[tag + '/var_param'], g[tag + '/var_param1']
    x = fam.sample(vp, 64, base=base)
    assert isinstance(x, np.ndarray)
    assert relerr(x, g[tag + '/sample']) < 1e-12
    assert relerr(fam.log_density(vp, x), g[tag + '/log_density']) < TOL64
    assert relerr(fam.log_density(vp, x[0]), g[tag + '/log_density_1d']) < TOL64
    assert relerr(fam.entropy(vp), g[tag + '/entropy']) < TOL64
    assert relerr(fam.init_param(), g[tag + '/init_param']) < 1e-15
    if fam.supports_kl:
        assert relerr(fam.kl(vp, vp1), g[tag + '/kl']) < TOL64
    else:
        with pytest.raises(NotImplementedError):
            fam.kl(vp, vp1)
    mean, cov = fam.mean_and_cov(vp)
    assert relerr(mean, g[tag + '/mean']) < TOL64 and relerr(cov, g[tag + '/cov']) < TOL64
    for p in (2, 4):
        key = tag + '/moment%d' % p
        if key in g:
            assert relerr(fam.pth_moment(vp, p), g[key]) < TOL64
        else:
            with pytest.raises(ValueError):
                fam.pth_moment(vp, p)
    # tensors in -> tensors out
    xt = fam.sample(torch.as_tensor(vp, device='cuda'), 5)
    assert isinstance(xt, torch.Tensor) and xt.shape == (5, d)


def _torch_models(vb):
    """The golden problems as viabel_b200 models."""
    m = {}
    for name, (N, d, seed, cls) in {'logistic_d4': (60, 4, 11, 'LogisticRegression'),
                                    'logistic_d10': (1000, 10, 12, 'LogisticRegression'),
                                    'probit_d6': (200, 6, 13, 'ProbitRegression')}.items():
        X, y, _ = logistic_problem(N, d, seed=seed)
        m[name] = getattr(vb, cls)(X, y, prior_scale=10.0)
    mean, sd = target_params(5, seed=14)
    mt, st = torch.as_tensor(mean, device='cuda'), torch.as_tensor(sd, device='cuda')

    def gauss(x):
        z = (x - mt) / st
        return (-0.5 * z * z - torch.log(st) - 0.5 * np.log(2 * np.pi)).sum(dim=1)

    def gauss_grad(x):
        return -(x - mt) / (st * st)

    m['gauss_d5'] = vb.Model(gauss, gauss_grad)

    def student(x):
        df = 10.0
        z = (x - mt) / st
        c = math.lgamma(0.5 * (df + 1)) - math.lgamma(0.5 * df) - 0.5 * np.log(df * np.pi)
        return (c - 0.5 * (df + 1) * torch.log1p(z * z / df) - torch.log(st)).sum(dim=1)

    m['student_d5'] = vb.Model(student)        # gradient by torch autograd

    hp = hier_problem(G=3, p=2, n_per=7, seed=15)
    Xh = torch.as_tensor(hp['X'], device='cuda')
    yh = torch.as_tensor(hp['y'], device='cuda')
    grp = torch.as_tensor(hp['group'], device='cuda')

    def hier(theta):
        G, p = 3, 2
        S = theta.shape[0]
        beta = theta[:, :G * p].reshape(S, G, p)
        mm = theta[:, G * p:G * p + p]
        ltau, lsig = theta[:, -2], theta[:, -1]
        pred = torch.einsum('np,snp->sn', Xh, beta[:, grp, :])
        c = 0.5 * np.log(2 * np.pi)
        res = (yh[None, :] - pred) / torch.exp(lsig)[:, None]
        lp = (-0.5 * res ** 2).sum(dim=1) - Xh.shape[0] * (lsig + c)
        db = (beta - mm[:, None, :]) / torch.exp(ltau)[:, None, None]
        lp = lp + (-0.5 * db ** 2).sum(dim=(1, 2)) - G * p * (ltau + c)
        lp = lp + (-0.5 * (mm / 10.0) ** 2).sum(dim=1) - p * (np.log(10.0) + c)
        return lp - 0.5 * ltau ** 2 - c - 0.5 * lsig ** 2 - c

    m['hier_G3p2'] = vb.Model(hier)
    return m


def test_objectives_golden(vb, golden):
    g = golden('objectives')
    models = _torch_models(vb)
    tags = sorted({k.rsplit('/', 1)[0] for k in g if k.endswith('/value')})
    checked = 0
    for tag in tags:
        mname, fam, point, oname = tag.split('/')
        kind, df = fam.split('_df')
        if mname not in models:
            continue
        vp = g[tag + '/var_param']
        if kind == 'mvt':
            base = (g[tag + '/chi2'], g[tag + '/z'])
            d = base[1].shape[1]
            approx = vb.MultivariateT(d, float(df))
            nS = base[1].shape[0]
        else:
            base = g[tag + '/base']
            d = base.shape[1]
            approx = vb.MFGaussian(d) if kind == 'mfg' else vb.MFStudentT(d, float(df))
            nS = base.shape[0]
        if oname == 'ekl':
            obj = vb.ExclusiveKL(approx, models[mname], nS)
        elif oname == 'ekl_path':
            obj = vb.ExclusiveKL(approx, models[mname], nS, use_path_deriv=True)
        else:
            obj = vb.AlphaDivergence(approx, models[mname], nS, float(oname[5:]))
        value, grad = obj(vp, base=base)
        assert isinstance(grad, np.ndarray)
        assert relerr(value, g[tag + '/value']) < TOL64, (tag, 'value')
        assert relerr(grad, g[tag + '/grad']) < TOL64, (tag, 'grad')
        checked += 1
    assert checked >= 100


@pytest.mark.parametrize('N,d,S', [(1000, 10, 10), (1003, 13, 7), (20000, 512, 256), (4100, 130, 300),
                                   (33, 1024, 40), (9, 2048, 3)])
@pytest.mark.parametrize('link', ['logistic', 'probit'])
def test_glm_sweep_vs_oracle(vb, vo, N, d, S, link):
    """Ragged shapes, every tile size (BM = 32 / 16 / 8) and multi-chunk S."""
    X, y, beta = logistic_problem(N, d, seed=N + d)
    rs = np.random.RandomState(S)
    base = rs.randn(S, d)
    vp = np.concatenate([beta + 0.01 * rs.randn(d), -2.0 + 0.1 * rs.randn(d)])
    cls = vb.LogisticRegression if link == 'logistic' else vb.ProbitRegression
    model = cls(X, y, prior_scale=10.0)
    fn = vo.logistic_logp_grad if link == 'logistic' else vo.probit_logp_grad
    oracle_model = lambda th: fn(th, X, y, 10.0)
    for fam_name, approx in (('gaussian', vb.MFGaussian(d)), ('student', vb.MFStudentT(d, 7.0))):
        df = 7.0
        v, gr = vb.ExclusiveKL(approx, model, S)(vp, base=base)
        v0, g0, f0 = vo.exclusive_kl_meanfield(vp, base, oracle_model, fam_name, df)
        assert relerr(v, v0) < TOL64 and relerr(gr, g0) < TOL64
        v, gr = vb.AlphaDivergence(approx, model, S, 2.0)(vp, base=base)
        v0, g0, _ = vo.alpha_divergence_meanfield(vp, base, oracle_model, 2.0, fam_name, df)
        assert relerr(v, v0) < TOL64 and relerr(gr, g0) < TOL64
    # the model's own __call__ (log density only)
    theta = vo.mfg_sample(vp, base)
    assert relerr(model(theta), oracle_model(theta)[0]) < TOL64


def test_glm_edge_cases(vb, vo):
    X, y, _ = logistic_problem(50, 6, seed=1)
    model = vb.LogisticRegression(X, y)
    approx = vb.MFGaussian(6)
    obj = vb.ExclusiveKL(approx, model, 4)
    with pytest.raises(ValueError):
        obj(np.zeros(5))                        # wrong var_param length
    with pytest.raises(ValueError):
        vb.LogisticRegression(X, y[:-1])
    with pytest.raises(ValueError):
        vb.MFStudentT(3, 2.0)
    with pytest.raises(ValueError):
        vb.ExclusiveKL(approx, model, 4, hessian_approx_method='invalid method')
    # extreme logits: saturated sigmoid / softplus stay finite (init sigma = e^2, large |z|)
    v, gr = obj(approx.init_param() * 5)
    assert np.isfinite(v) and np.all(np.isfinite(gr))
    # single observation, single sample
    m1 = vb.LogisticRegression(X[:1], y[:1])
    base = np.random.RandomState(0).randn(1, 6)
    v, gr = vb.ExclusiveKL(approx, m1, 1)(approx.init_param(), base=base)
    v0, g0, _ = vo.exclusive_kl_meanfield(approx.init_param(), base,
                                          lambda th: vo.logistic_logp_grad(th, X[:1], y[:1], 10.0))
    assert relerr(v, v0) < TOL64 and relerr(gr, g0) < TOL64


def test_optimizer_steps_golden(vb, golden):
    g = golden('optimizers')
    grads = g['grads']
    for name, mk in (('rmsprop', lambda: vb.RMSProp(0.01)), ('adam', lambda: vb.Adam(0.01)),
                     ('rmsprop_b', lambda: vb.RMSProp(0.01, beta=0.5, jitter=1e-6)),
                     ('adam_b', lambda: vb.Adam(0.01, beta1=0.7, beta2=0.9, jitter=1e-6))):
        host_opt, dev_opt = mk(), mk()
        vp = torch.zeros(grads.shape[1], dtype=torch.float64, device='cuda')
        for i, gr in enumerate(grads):
            assert relerr(host_opt.descent_direction(gr.copy()), g[name + '/dirs'][i]) < 1e-13
            before = vp.clone()
            d = dev_opt._fused_step(vp, torch.as_tensor(gr, device='cuda'), True)
            assert relerr(d.cpu().numpy(), g[name + '/dirs'][i]) < 1e-13, (name, i)
            assert relerr((before - vp).cpu().numpy(), 0.01 * g[name + '/dirs'][i]) < 1e-12


def test_optimize_loop_golden(vb, golden):
    g = golden('optimizers')

    class Quad(vb.VariationalObjective):
        def __init__(self):
            pass

        def _update_objective_and_grad(self):
            pass

        def __call__(self, vp):
            return 0.5 * (vp ** 2).sum(), vp.clone()

    for name, mk in (('rmsprop', lambda: vb.RMSProp(0.1)), ('adam', lambda: vb.Adam(0.1))):
        opt = mk()
        opt.progress = False
        res = opt.optimize(25, Quad(), np.array([1.0, -2.0, 0.5]))
        assert relerr(res['opt_param'], g[name + '/opt_param']) < 1e-12
        assert relerr(res['value_history'], g[name + '/value_history']) < 1e-12
        assert res['variational_param_history'].shape == g[name + '/param_history'].shape
        assert relerr(res['variational_param_history'], g[name + '/param_history']) < 1e-12


def test_bbvi_style_fit_converges(vb):
    """tests/test_objectives.py:11-32 scenario: RMSProp(0.1), 1000 iterations, S=100, MFStudentT(2,100)
    on a 2-D Gaussian target; recover mean / sd to 1 decimal."""
    mean = torch.tensor([1., -1.], dtype=torch.float64, device='cuda')
    sd = torch.tensor([2., 5.], dtype=torch.float64, device='cuda')

    def log_p(x):
        z = (x - mean) / sd
        return (-0.5 * z * z - torch.log(sd) - 0.5 * np.log(2 * np.pi)).sum(dim=1)

    fits = {}
    for name, make in (('ekl', lambda a: vb.ExclusiveKL(a, log_p, 100)),
                       ('path', lambda a: vb.ExclusiveKL(a, log_p, 100, use_path_deriv=True))):
        np.random.seed(851)
        approx = vb.MFStudentT(2, 100)
        opt = vb.RMSProp(0.1)
        opt.progress = False
        res = opt.optimize(1000, make(approx), np.array([0, 0, 1, 1], dtype=np.float32))
        est_mean, est_cov = approx.mean_and_cov(res['opt_param'])
        np.testing.assert_almost_equal(est_mean, [1., -1.], decimal=1)
        np.testing.assert_almost_equal(np.sqrt(np.diag(est_cov)), [2., 5.], decimal=1)
        fits[name] = res['opt_param']
    # AlphaDivergence (tests/test_objectives.py:90-91).  The reference's S=100 CUBO estimator
    # collapses (sigma -> 0) from the cold start for roughly one seed in three -- the numpy oracle
    # shows the same -- so the scenario is started from the ELBO fit, where it is stable.
    np.random.seed(851)
    approx = vb.MFStudentT(2, 100)
    opt = vb.RMSProp(0.02)
    opt.progress = False
    res = opt.optimize(600, vb.AlphaDivergence(approx, log_p, 100, 2), fits['ekl'])
    est_mean, est_cov = approx.mean_and_cov(res['opt_param'])
    np.testing.assert_allclose(est_mean, [1., -1.], atol=0.35)
    np.testing.assert_allclose(np.sqrt(np.diag(est_cov)), [2., 5.], rtol=0.12)


def test_builtin_target_and_hier_plugins(vb, vo):
    rs = np.random.RandomState(4)
    mean, sd = target_params(6, seed=30)
    theta = rs.randn(9, 6)
    for plug, ref in ((vb.GaussianTarget(mean, sd), lambda t: vo.gauss_target_logp_grad(t, mean, sd)),
                      (vb.StudentTTarget(mean, sd, 10.0), lambda t: vo.student_target_logp_grad(t, mean, sd, 10.0))):
        lp, g = plug.logp_and_grad(torch.as_tensor(theta, device='cuda'))
        lp0, g0 = ref(theta)
        assert relerr(lp.cpu().numpy(), lp0) < TOL64 and relerr(g.cpu().numpy(), g0) < TOL64
        assert relerr(plug(theta), lp0) < TOL64
    hp = hier_problem(G=5, p=3, n_per=11, seed=31)
    model = vb.HierarchicalLinearRegression(hp['X'], hp['y'], hp['group'], 5)
    theta = 0.3 * rs.randn(7, model.dim)
    lp, g = model.logp_and_grad(torch.as_tensor(theta, device='cuda'))
    lp0, g0 = vo.hier_linear_logp_grad(theta, hp['X'], hp['y'], hp['group'], 5, 3)
    assert relerr(lp.cpu().numpy(), lp0) < TOL64 and relerr(g.cpu().numpy(), g0) < TOL64
    # full-rank t + AlphaDivergence on it (BASELINE configs[3] at a small size) against the oracle
    approx = vb.MultivariateT(model.dim, 100)
    vp = approx.init_param()
    vp[model.dim:] *= 0.0
    vp[model.dim:] += vo.mvt_pack(np.zeros(model.dim), 0.05 * np.eye(model.dim) + 0.01)[model.dim:]
    chi2, z = rs.chisquare(100, 16), rs.randn(16, model.dim)
    om = lambda th: vo.hier_linear_logp_grad(th, hp['X'], hp['y'], hp['group'], 5, 3)
    v, gr = vb.AlphaDivergence(approx, model, 16, 2.0)(vp, base=(chi2, z))
    v0, g0, _ = vo.alpha_divergence_mvt(vp, chi2, z, om, 100.0, 2.0)
    assert relerr(v, v0) < TOL64 and relerr(gr, g0) < TOL64
    v, gr = vb.ExclusiveKL(approx, model, 16)(vp, base=(chi2, z))
    v0, g0, _ = vo.exclusive_kl_mvt(vp, chi2, z, om, 100.0)
    assert relerr(v, v0) < TOL64 and relerr(gr, g0) < TOL64


def test_dis_inclusive_kl(vb):
    """DISInclusiveKL (objectives.py:280-416): score-function gradient at the fixed samples against
    torch autograd, ESS bisection invariants, and the reference's convergence scenario
    (tests/test_objectives.py:82-87)."""
    mean = torch.tensor([1., -1.], dtype=torch.float64, device='cuda')
    sd = torch.tensor([2., 5.], dtype=torch.float64, device='cuda')
    target = vb.GaussianTarget(mean, sd)
    for fam in (vb.MFGaussian(2), vb.MFStudentT(2, 100)):
        obj = vb.DISInclusiveKL(fam, target, 200, ess_target=50, temper_prior=vb.MFGaussian(2),
                                temper_prior_params=np.array([0., 0., 1., 1.]), use_resampling=False)
        vp = torch.tensor([0.3, -0.2, 0.5, 1.2], dtype=torch.float64, device='cuda')
        value, grad = obj(vp)
        xs, w = obj._state_samples, obj._state_w
        assert 0.0 <= obj._eps <= 1.0 and bool(torch.isfinite(w).all())
        vpt = vp.clone().requires_grad_(True)
        z = (xs - vpt[:2]) / torch.exp(vpt[2:])
        if isinstance(fam, vb.MFGaussian):
            logq = (-0.5 * z * z - vpt[2:] - 0.5 * np.log(2 * np.pi)).sum(dim=1)
        else:
            df = 100.0
            c = math.lgamma(0.5 * (df + 1)) - math.lgamma(0.5 * df) - 0.5 * math.log(df * math.pi)
            logq = (c - 0.5 * (df + 1) * torch.log1p(z * z / df) - vpt[2:]).sum(dim=1)
        ref = -(w / 200 * logq).sum()
        (gref,) = torch.autograd.grad(ref, vpt)
        assert relerr(float(value), float(ref)) < 1e-12
        assert relerr(grad.cpu().numpy(), gref.cpu().numpy()) < 1e-12
    np.random.seed(851)
    approx = vb.MFStudentT(2, 100)
    obj = vb.DISInclusiveKL(approx, target, 100, ess_target=50, temper_prior=vb.MFGaussian(2),
                            temper_prior_params=np.concatenate([[0] * 2, [1] * 2]))
    opt = vb.RMSProp(0.1)
    opt.progress = False
    res = opt.optimize(1000, obj, np.array([0, 0, 1, 1], dtype=np.float32))
    est_mean, est_cov = approx.mean_and_cov(res['opt_param'])
    np.testing.assert_allclose(est_mean, [1., -1.], atol=0.5)
    np.testing.assert_allclose(np.sqrt(np.diag(est_cov)), [2., 5.], rtol=0.25)


@pytest.mark.parametrize('G,p,n_per,S', [(1, 5, 40, 9), (3, 20, 30, 24), (7, 31, 40, 40)])
def test_full_rank_path_vs_oracle(vb, vo, G, p, n_per, S):
    """MultivariateT + hierarchical linear regression (BASELINE configs[3]) at d = 12, 82 and 250 against the oracle:
    sample, log density, mean / covariance / moments, ExclusiveKL and AlphaDivergence(alpha = 2) value and gradient,
    all to 1e-10 -- unpack, Sigma, the eigenbasis reparameterisation and every cotangent GEMM are the package's own
    float64 tensor-core kernels (csrc/mvt.cu, gemm_f64.cu, hier.cu); only eigh is a library call."""
    hp = hier_problem(G=G, p=p, n_per=n_per, seed=31 + G)
    model = vb.HierarchicalLinearRegression(hp['X'], hp['y'], hp['group'], G)
    d = model.dim
    rs = np.random.RandomState(d)
    om = lambda th: vo.hier_linear_logp_grad(th, hp['X'], hp['y'], hp['group'], G, p)
    theta = 0.3 * rs.randn(5, d)
    lp, g = model.logp_and_grad(torch.as_tensor(theta, device='cuda'))
    lp0, g0 = om(theta)
    assert relerr(lp.cpu().numpy(), lp0) < TOL64 and relerr(g.cpu().numpy(), g0) < TOL64
    approx = vb.MultivariateT(d, 100)
    # eigenvalues around 1: the reference's entropy is 0.5 * log(det(Sigma)) (approximations.py:354), and det()
    # under- / overflows away from unit scale at d = 250 (parity where the reference's det is finite)
    B = 0.3 * rs.randn(d, d) / np.sqrt(d)
    Sigma = 0.9 * np.eye(d) + B @ B.T
    mu = 0.2 * rs.randn(d)
    vp = vo.mvt_pack(mu, Sigma)
    chi2, z = rs.chisquare(100, S), rs.randn(S, d)
    x = approx.sample(vp, S, base=(chi2, z))
    x0 = vo.mvt_sample(vp, chi2, z, 100.0)
    assert relerr(x, x0) < TOL64
    assert relerr(approx.log_density(vp, x0), vo.mvt_log_density(vp, x0, 100.0)) < TOL64
    mean, cov = approx.mean_and_cov(vp)
    mean0, cov0 = vo.mvt_mean_and_cov(vp, d, 100.0)
    assert relerr(mean, mean0) < 1e-14 and relerr(cov, cov0) < 1e-12
    for pm in (2, 4):
        assert relerr(approx.pth_moment(vp, pm), vo.mvt_pth_moment(vp, d, 100.0, pm)) < TOL64
    v, gr = vb.ExclusiveKL(approx, model, S)(vp, base=(chi2, z))
    v0, g0, _ = vo.exclusive_kl_mvt(vp, chi2, z, om, 100.0)
    assert relerr(v, v0) < TOL64 and relerr(gr, g0) < TOL64
    v, gr = vb.AlphaDivergence(approx, model, S, 2.0)(vp, base=(chi2, z))
    v0, g0, _ = vo.alpha_divergence_mvt(vp, chi2, z, om, 100.0, 2.0)
    assert relerr(v, v0) < TOL64 and relerr(gr, g0) < TOL64


@pytest.mark.parametrize('ta,tb,M,N,K', [(0, 0, 70, 130, 33), (1, 0, 64, 64, 64), (0, 1, 5, 300, 129), (1, 1, 257, 31, 1),
                                         (0, 0, 1, 200, 200)])
def test_gemm_f64_epilogues(vb, ta, tb, M, N, K):
    """vb_gemm_f64 (float64 DMMA GEMM) with every fused scaling against numpy, ragged shapes."""
    rs = np.random.RandomState(M + N + K)
    A = rs.randn(K, M) if ta else rs.randn(M, K)
    Bm = rs.randn(N, K) if tb else rs.randn(K, N)
    ks, rsc, bias = rs.rand(K) + 0.5, rs.randn(M), rs.randn(N)
    dm, dn = rs.rand(M) + 0.5, rs.rand(N) + 0.5
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device='cuda')
    C = torch.empty(M, N, dtype=torch.float64, device='cuda')
    lib, ptr = vb._lib.lib, vb._lib.ptr
    Ad, Bd, ksd, rsd, bd, dmd, dnd = t(A), t(Bm), t(ks), t(rsc), t(bias), t(dm), t(dn)
    vb._lib.check(lib.vb_gemm_f64(ta, tb, M, N, K, 0.7, ptr(Ad), Ad.shape[1], ptr(Bd), Bd.shape[1], ptr(C), N, ptr(ksd), ptr(rsd),
                                  ptr(bd), ptr(dmd), ptr(dnd), vb._lib.stream()))
    Aop = A.T if ta else A
    Bop = Bm.T if tb else Bm
    ref = 0.7 * ((Aop * ks) @ Bop) * rsc[:, None] / (dm[:, None] + dn[None, :]) + bias
    assert relerr(C.cpu().numpy(), ref) < 1e-13
    vb._lib.check(lib.vb_gemm_f64(ta, tb, M, N, K, 1.0, ptr(Ad), Ad.shape[1], ptr(Bd), Bd.shape[1], ptr(C), N, None, None, None,
                                  None, None, vb._lib.stream()))
    assert relerr(C.cpu().numpy(), Aop @ Bop) < 1e-13


def test_stan_model_adapter(vb, vo):
    """StanModel around a fit-like object (host log_prob / grad_log_prob, one row at a time): values, per-sample
    gradients and an ExclusiveKL evaluation against the oracle on the same Gaussian target."""
    d, S = 6, 40
    rs = np.random.RandomState(12)
    mean, sd = rs.randn(d), 0.5 + rs.rand(d)

    class Fit:
        def log_prob(self, x):
            return float(np.sum(-0.5 * ((np.asarray(x) - mean) / sd) ** 2 - np.log(sd) - 0.5 * np.log(2 * np.pi)))

        def grad_log_prob(self, x):
            return -(np.asarray(x) - mean) / sd ** 2

        def constrain_pars(self, x):
            return x

    model = vb.StanModel(Fit())
    theta = rs.randn(S, d)
    ref = np.array([Fit().log_prob(t) for t in theta])
    assert relerr(model(theta), ref) < TOL64
    assert abs(model(theta[0]) - ref[0]) < 1e-12
    lp, g = model.logp_and_grad(torch.as_tensor(theta, device='cuda'))
    assert relerr(lp.cpu().numpy(), ref) < TOL64
    assert relerr(g.cpu().numpy(), -(theta - mean) / sd ** 2) < TOL64
    approx = vb.MFGaussian(d)
    base = rs.randn(S, d)
    vp = np.concatenate([0.3 * rs.randn(d), -0.5 + 0.1 * rs.randn(d)])
    v, gr = vb.ExclusiveKL(approx, model, S)(vp, base=base)
    v0, g0, _ = vo.exclusive_kl_meanfield(vp, base, lambda th: vo.gauss_target_logp_grad(th, mean, sd))
    assert relerr(v, v0) < TOL64 and relerr(gr, g0) < TOL64
