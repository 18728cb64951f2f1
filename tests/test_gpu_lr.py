"""GPU parity for LRGaussian (approximations.py:610-731) against the unmodified reference
(tests/golden/lr_gaussian.npz, oracle/make_golden.py: gen_lr_gaussian) and the oracle at larger sizes."""
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


@pytest.mark.parametrize('d,k', [(3, 0), (3, 1), (8, 3), (6, 6)])
def test_lr_family_golden(vb, golden, d, k):
    g = golden('lr_gaussian')
    t = 'lr_d%d_k%d' % (d, k)
    fam = vb.LRGaussian(d, k=k)
    vp, vp1 = g[t + '/var_param'], g[t + '/var_param1']
    assert fam.var_param_dim == vp.size and fam.supports_kl and fam.supports_entropy
    x = fam.sample(vp, 40, base=(g[t + '/z'], g[t + '/eps']))
    assert relerr(x, g[t + '/sample']) < 1e-13
    assert relerr(fam.log_density(vp, x), g[t + '/log_density']) < TOL64
    assert relerr(fam.log_density(vp, x[0]), g[t + '/log_density_1d']) < TOL64
    assert relerr(fam.entropy(vp), g[t + '/entropy']) < TOL64
    assert relerr(fam.kl(vp, vp1), g[t + '/kl']) < TOL64
    mean, cov = fam.mean_and_cov(vp)
    assert relerr(mean, g[t + '/mean']) < 1e-14 and relerr(cov, g[t + '/cov']) < 1e-13
    for p in (2, 4):
        assert relerr(fam.pth_moment(vp, p), g[t + '/moment%d' % p]) < TOL64
    init = fam.init_param()
    assert init.shape == (fam.var_param_dim,) and relerr(init[:2 * d], g[t + '/init_head']) == 0
    # native draws: z first, then eps, consecutive slices of one Philox stream; explicit seed restarts it
    a = fam.sample(vp, 5, seed=9)
    b = fam.sample(vp, 5, seed=9)
    assert np.array_equal(a, b)
    z, eps = fam.last_base
    assert tuple(z.shape) == (5, k) and tuple(eps.shape) == (5, d)


def test_lr_objectives_golden(vb, golden):
    g = golden('lr_gaussian')
    X, y, _ = logistic_problem(60, 4, seed=11)
    mean, sd = target_params(5, seed=14)
    models = {'logistic_d4': vb.LogisticRegression(X, y, prior_scale=10.0), 'gauss_d5': vb.GaussianTarget(mean, sd)}
    n = 0
    for key in [q for q in g if q.startswith('obj/') and q.endswith('/value')]:
        t = key[:-len('/value')]
        _, mname, kk, oname = t.split('/')
        d = 4 if mname == 'logistic_d4' else 5
        fam = vb.LRGaussian(d, k=int(kk[1:]))
        base = (g[t + '/z'], g[t + '/eps'])
        if oname == 'alpha2':
            obj = vb.AlphaDivergence(fam, models[mname], 7, 2.0)
        else:
            obj = vb.ExclusiveKL(fam, models[mname], 7, use_path_deriv=(oname == 'ekl_path'))
        v, gr = obj(g[t + '/var_param'], base=base)
        assert relerr(v, g[t + '/value']) < TOL64, t
        assert relerr(gr, g[t + '/grad']) < 1e-9, t
        n += 1
    assert n == 12


def test_lr_vs_oracle_larger_and_fit(vb, vo):
    d, k, S = 40, 5, 30
    X, y, beta = logistic_problem(3000, d, seed=77)
    rs = np.random.RandomState(1)
    vp = np.concatenate([0.5 * beta, -1.0 + 0.1 * rs.randn(d), 0.1 * rs.randn(d * k)])
    z, eps = rs.randn(S, k), rs.randn(S, d)
    model = vb.LogisticRegression(X, y)
    om = lambda th: vo.logistic_logp_grad(th, X, y, 10.0)
    fam = vb.LRGaussian(d, k=k)
    for kind, obj in (('ekl', vb.ExclusiveKL(fam, model, S)), ('ekl_path', vb.ExclusiveKL(fam, model, S, use_path_deriv=True)),
                      ('alpha', vb.AlphaDivergence(fam, model, S, 2.0))):
        v, gr = obj(vp, base=(z, eps))
        v0, g0 = vo.lr_objective(vp, z, eps, om, kind, 2.0)
        assert relerr(v, v0) < TOL64 and relerr(gr, g0) < 1e-9, kind
    vp1 = vp + 0.05 * rs.randn(vp.size)
    assert relerr(fam.kl(vp, vp1), vo.lr_kl(vp, vp1, d, k)) < TOL64
    # a short fit moves towards the mode: the ELBO improves
    opt = vb.RMSProp(0.05)
    opt.progress = False
    res = opt.optimize(150, vb.ExclusiveKL(fam, model, S), fam.init_param())
    assert np.mean(res['value_history'][-20:]) < np.mean(res['value_history'][:20])
