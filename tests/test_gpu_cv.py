"""GPU parity for the control-variate ExclusiveKL estimators (objectives.py:170-273) and the kernels under them:
vb_glm_point_f64 (gradient / Hessian-vector products / Hessian at one point) and vb_sample_moments_f64 (column
means, central power sums, covariance on the FP64 tensor pipe).  Goldens come from the unmodified reference run
through the autograd stand-in (oracle/make_golden.py: gen_objectives_cv)."""
import numpy as np
import pytest
import torch

from conftest import relerr
from _problems import logistic_problem, target_params

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


def _product_model(vb, name):
    if name.startswith('logistic'):
        N, d, seed = (60, 4, 11) if name == 'logistic_d4' else (1000, 10, 12)
        X, y, _ = logistic_problem(N, d, seed=seed)
        return vb.LogisticRegression(X, y, prior_scale=10.0)
    if name == 'probit_d6':
        X, y, _ = logistic_problem(200, 6, seed=13)
        return vb.ProbitRegression(X, y, prior_scale=10.0)
    mean, sd = target_params(5, seed=14)
    if name == 'gauss_d5':
        return vb.GaussianTarget(mean, sd)
    return vb.StudentTTarget(mean, sd, 10.0)


@pytest.mark.parametrize('mname', ['logistic_d4', 'logistic_d10', 'probit_d6', 'gauss_d5', 'student_d5'])
def test_control_variate_objectives_golden(vb, golden, mname):
    g = golden('objectives_cv')
    model = _product_model(vb, mname)
    n = 0
    for key in [k for k in g if k.startswith(mname + '/') and k.endswith('/value')]:
        tag = key[:-len('/value')]
        _, fam, method, mode = tag.split('/')
        base = g[tag + '/base']
        d = base.shape[1]
        approx = vb.MFGaussian(d) if fam.startswith('mfg') else vb.MFStudentT(d, 8)
        obj = vb.ExclusiveKL(approx, model, base.shape[0], use_path_deriv=(mode == 'path'),
                             hessian_approx_method=method)
        v, gr = obj(g[tag + '/var_param'], base=base)
        assert relerr(v, g[tag + '/value']) < TOL64, tag
        assert relerr(gr, g[tag + '/grad']) < TOL64, tag
        n += 1
    assert n == 16


def test_control_variates_generic_model_autograd(vb, golden):
    """A user Model(log_density) on CUDA tensors: gradient / HVP / Hessian at the mean come from torch autograd."""
    g = golden('objectives_cv')
    mean, sd = target_params(5, seed=14)
    mt, st = torch.as_tensor(mean, device='cuda'), torch.as_tensor(sd, device='cuda')

    def log_p(x):
        z = (x - mt) / st
        return (-0.5 * z * z - torch.log(st) - 0.5 * np.log(2 * np.pi)).sum(dim=1)

    for method in ('full', 'mean_only', 'loo_diag_approx', 'loo_direct_approx'):
        tag = 'gauss_d5/mfg_dfNone/%s/plain' % method
        base = g[tag + '/base']
        obj = vb.ExclusiveKL(vb.MFGaussian(5), log_p, base.shape[0], hessian_approx_method=method)
        v, gr = obj(g[tag + '/var_param'], base=base)
        assert relerr(v, g[tag + '/value']) < TOL64 and relerr(gr, g[tag + '/grad']) < TOL64, tag


@pytest.mark.parametrize('N,d,link', [(5000, 300, 'logistic'), (777, 13, 'probit'), (40000, 512, 'logistic'),
                                      (3000, 2048, 'probit'), (1, 5, 'logistic')])
def test_glm_point_derivatives_vs_oracle(vb, vo, N, d, link):
    X, y, beta = logistic_problem(N, d, seed=N + d)
    rs = np.random.RandomState(d)
    m = 0.7 * beta + 0.02 * rs.randn(d)
    V = rs.randn(3, d)
    model = (vb.LogisticRegression if link == 'logistic' else vb.ProbitRegression)(X, y, prior_scale=3.0)
    want_h = d <= 512
    gdev, HV, H = model.point_derivatives(torch.as_tensor(m, device='cuda'), torch.as_tensor(V, device='cuda'), want_h)
    fn = vo.logistic_logp_grad if link == 'logistic' else vo.probit_logp_grad
    hs = vo.logistic_hessian if link == 'logistic' else vo.probit_hessian
    g0 = fn(m[None, :], X, y, 3.0)[1][0]
    H0 = hs(m, X, y, 3.0)
    assert relerr(gdev.cpu().numpy(), g0) < TOL64
    assert relerr(HV.cpu().numpy(), V @ H0) < TOL64
    if want_h:
        assert relerr(H.cpu().numpy(), H0) < TOL64
        assert np.array_equal(H.cpu().numpy(), H.cpu().numpy().T)
    # more than 8 vectors: split into passes
    V12 = rs.randn(12, d)
    _, HV12, _ = model.point_derivatives(torch.as_tensor(m, device='cuda'), torch.as_tensor(V12, device='cuda'))
    assert relerr(HV12.cpu().numpy(), V12 @ H0) < TOL64


@pytest.mark.parametrize('n,d', [(20000, 3), (1000, 256), (50001, 70), (7, 5), (2, 1), (300000, 130)])
def test_sample_moments_vs_numpy(vb, n, d):
    rs = np.random.RandomState(n + d)
    x = rs.randn(n, d) * np.exp(0.5 * rs.randn(d)) + 3.0 * rs.randn(d)
    from viabel_b200.diagnostics import sample_moments
    mean, m2, m4, cov = sample_moments(torch.as_tensor(x, device='cuda'), want_cov=True)
    xc = x - x.mean(axis=0)
    assert relerr(mean.cpu().numpy(), x.mean(axis=0)) < 1e-12
    assert relerr(m2.cpu().numpy(), (xc ** 2).sum(axis=0)) < 1e-12
    assert relerr(m4.cpu().numpy(), (xc ** 4).sum(axis=0)) < 1e-12
    ref = np.atleast_2d(np.cov(x.T))
    assert relerr(cov.cpu().numpy(), ref) < 1e-11
    assert np.array_equal(cov.cpu().numpy(), cov.cpu().numpy().T)


def test_diagnostics_sample_branch_uses_the_kernels(vb, vo):
    """wasserstein_bounds / all_diagnostics with samples (no moment_bound_fn, no q_var): the sample-moment branch
    (diagnostics.py:137-141, :58-59) against the oracle."""
    rs = np.random.RandomState(846)
    n, d = 30000, 4
    samples = rs.randn(n, d) * np.array([1.0, 2.0, 0.5, 1.5]) + np.array([0.3, -1.0, 2.0, 0.0])
    lw = -0.1 * np.sum(samples ** 2, axis=1) + 0.05 * rs.randn(n)
    res = vb.all_diagnostics(lw, samples=samples)
    ref = vo.all_diagnostics(lw, samples=samples)
    for k in ('d2', 'W1', 'W2', 'mean_error', 'std_error', 'cov_error', 'log_norm_bound'):
        assert relerr(res[k], ref[k]) < TOL64, k
    # 1-D samples are promoted to a column
    res1 = vb.wasserstein_bounds(0.3, samples=samples[:, 0])
    ref1 = vo.wasserstein_bounds(0.3, samples=samples[:, 0])
    assert relerr(res1['W1'], ref1['W1']) < TOL64 and relerr(res1['W2'], ref1['W2']) < TOL64
