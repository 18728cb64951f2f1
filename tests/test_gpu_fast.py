"""GPU parity for the tensor-core fast path (tcgen05 + TMA, fp16 hi/lo operand splits).
Tolerance stated by BASELINE.json north_star for the FP32/TF32 fast path: 1e-4 relative on the
ELBO value and (norm-wise) on the gradient.  The float64 oracle is the reference."""
import numpy as np
import pytest
import torch

from conftest import relerr
from _problems import logistic_problem

pytestmark = pytest.mark.gpu
TOL_FAST = 1e-4


@pytest.fixture(scope='module')
def vb():
    import viabel_b200
    return viabel_b200


@pytest.fixture(scope='module')
def vo():
    from oracle import viabel_oracle
    return viabel_oracle


def _f16_exact(a):
    return a.astype(np.float16).astype(np.float64)


@pytest.mark.parametrize('N,d,S', [(1000, 512, 256), (20000, 512, 256), (1003, 13, 7), (4100, 130, 100),
                                   (128, 128, 64), (257, 1024, 256), (5, 2048, 3), (40000, 64, 256),
                                   (60000, 1024, 256), (45000, 2048, 64)])
def test_fast_sweep_vs_oracle(vb, vo, N, d, S):
    X, y, beta = logistic_problem(N, d, seed=N + d)
    rs = np.random.RandomState(S)
    base = _f16_exact(rs.randn(S, d))                      # fp16-exact draws (what quantize=2 generates)
    model = vb.LogisticRegression(X, y, prior_scale=10.0).enable_fast_path()
    oracle_model = lambda th: vo.logistic_logp_grad(th, X, y, 10.0)
    points = {'init': vo.mfg_init_param(d),
              'mid': np.concatenate([0.5 * beta, -1.0 * np.ones(d)]),
              'conv': np.concatenate([beta + 0.01 * rs.randn(d), -3.5 + 0.1 * rs.randn(d)])}
    approx = vb.MFGaussian(d)
    for name, vp in points.items():
        v, gr = vb.ExclusiveKL(approx, model, S)(vp, base=base)
        v0, g0, _ = vo.exclusive_kl_meanfield(vp, base, oracle_model)
        assert relerr(v, v0) < TOL_FAST, (name, 'value')
        assert relerr(gr, g0) < TOL_FAST, (name, 'grad')
        assert relerr(gr[:d], g0[:d]) < TOL_FAST and relerr(gr[d:], g0[d:]) < TOL_FAST, (name, 'grad parts')
    # weighted second pass (AlphaDivergence) and the forward-only model call
    vp = points['conv']
    v, gr = vb.AlphaDivergence(approx, model, S, 2.0)(vp, base=base)
    v0, g0, _ = vo.alpha_divergence_meanfield(vp, base, oracle_model, 2.0)
    assert relerr(v, v0) < TOL_FAST and relerr(gr, g0) < 5 * TOL_FAST     # weights exp(2 lw) amplify lw error
    theta = vo.mfg_sample(vp, base)
    assert relerr(model(theta), oracle_model(theta)[0]) < TOL_FAST
    # general float64 draws (not fp16-exact) still meet the tolerance
    base2 = rs.randn(S, d)
    v, gr = vb.ExclusiveKL(approx, model, S)(points['mid'], base=base2)
    v0, g0, _ = vo.exclusive_kl_meanfield(points['mid'], base2, oracle_model)
    assert relerr(v, v0) < TOL_FAST and relerr(gr, g0) < TOL_FAST


def test_fast_matches_f64_path_and_native_draws(vb):
    """Same model object, both paths, device-generated fp16-exact Philox draws."""
    N, d, S = 50000, 256, 256
    X, y, beta = logistic_problem(N, d, seed=9)
    model = vb.LogisticRegression(X, y)
    approx = vb.MFGaussian(d, seed=11)
    approx.quantize_draws = 2
    obj = vb.ExclusiveKL(approx, model, S)
    vp = np.concatenate([beta, -3.0 * np.ones(d)])
    v64, g64 = obj(vp)
    base = approx.last_base
    assert torch.equal(base, base.to(torch.float16).to(torch.float64))
    model.enable_fast_path()
    vf, gf = obj(vp, base=base)
    assert relerr(vf, v64) < TOL_FAST and relerr(gf, g64) < TOL_FAST
    model.path = 'f64'
    v2, g2 = obj(vp, base=base)
    assert relerr(v2, v64) < 1e-13 and relerr(g2, g64) < 1e-12


def test_fast_path_rejections(vb):
    X, y, _ = logistic_problem(300, 8, seed=2)
    with pytest.raises(NotImplementedError):
        vb.ProbitRegression(X, y).enable_fast_path()
    with pytest.raises(NotImplementedError):
        vb.LogisticRegression(X * 1e5, y).enable_fast_path()        # outside the fp16 operand range
    # more than 256 samples falls back to the float64 sweep transparently
    model = vb.LogisticRegression(X, y).enable_fast_path()
    approx = vb.MFGaussian(8)
    v, g = vb.ExclusiveKL(approx, model, 300)(approx.init_param())
    assert np.isfinite(v) and np.all(np.isfinite(g))
