"""NeuralNet / NVPFlow families (reference approximations.py:385-550) on the device against goldens
from the UNMODIFIED reference (oracle/make_golden.py::gen_flows; autograd replaced by the torch-backed shim):
forward functions, sampling with the prior's injected base draws, log density, and the two objectives the reference
can evaluate on a flow -- ExclusiveKL(use_path_deriv=True) and AlphaDivergence -- value and gradient at 1e-10."""
import numpy as np
import pytest
import torch

from conftest import relerr
from _problems import target_params

pytestmark = pytest.mark.gpu
TOL64 = 1e-10


@pytest.fixture(scope='module')
def vb():
    import viabel_b200
    return viabel_b200


def nvp_masks(dim, n_pairs):
    half, halfplus = dim // 2, dim - dim // 2
    m1 = np.hstack([[0] * half, [1] * halfplus])
    m2 = np.hstack([[1] * half, [0] * halfplus])
    return np.array(list(np.vstack([m1, m2])) * n_pairs)


@pytest.mark.parametrize('dim,hidden', [(1, 4), (3, 10), (6, 8)])
def test_neuralnet_forward_golden(vb, golden, dim, hidden):
    g = golden('flows')
    t = 'nn_d%d' % dim
    nn = vb.NeuralNet([[dim, hidden], [hidden, hidden], [hidden, dim]])
    assert nn.var_param_dim == g[t + '/flat'].size and nn.dim == dim
    y, ld = nn.forward(g[t + '/flat'], g[t + '/x'])                     # flat vector in PatternDict order
    assert relerr(y, g[t + '/y']) < TOL64 and relerr(ld, g[t + '/log_det_J']) < TOL64
    folded = {k: v.cpu().numpy() for k, v in nn.fold(torch.as_tensor(g[t + '/flat'], device='cuda')).items()}
    y2, _ = nn.forward(folded, g[t + '/x'])                             # folded dict, as the reference's tests pass it
    assert relerr(y2, g[t + '/y']) < TOL64
    s = nn.sample(folded, 5, base=g[t + '/x'][:5])
    assert relerr(s, g[t + '/y'][:5]) < TOL64
    with pytest.raises(NotImplementedError):
        nn.log_density(folded, g[t + '/x'])
    assert not nn.supports_entropy and not nn.supports_kl and not nn.supports_pth_moment(2)


@pytest.mark.parametrize('dim,hidden,pairs,df', [(1, 5, 2, None), (3, 10, 3, None), (6, 7, 2, 7)])
def test_nvpflow_golden(vb, golden, dim, hidden, pairs, df):
    g = golden('flows')
    t = 'nvp_d%d' % dim
    layers = [[dim, hidden], [hidden, dim]]
    prior = vb.MFGaussian(dim) if df is None else vb.MFStudentT(dim, df)
    fam = vb.NVPFlow(layers, layers, nvp_masks(dim, pairs), prior, g[t + '/prior_param'], dim)
    vp = g[t + '/var_param']
    assert fam.var_param_dim == vp.size and np.array_equal(fam.mask, g[t + '/mask'])
    x = fam.sample(vp, 11, base=g[t + '/base'])
    assert relerr(x, g[t + '/sample']) < TOL64
    assert relerr(fam.g(vp, g[t + '/z0']), g[t + '/sample']) < TOL64
    z, ld = fam.f(vp, g[t + '/sample'])
    assert relerr(z, g[t + '/f_z']) < TOL64 and relerr(ld, g[t + '/f_logdet']) < TOL64
    assert relerr(fam.log_density(vp, g[t + '/sample']), g[t + '/log_density']) < TOL64
    assert relerr(fam.log_density(vp, g[t + '/sample'][0]), g[t + '/log_density'][:1]) < TOL64     # 1-D x is promoted
    # device tensors in -> device tensors out
    xd = fam.sample(torch.as_tensor(vp, device='cuda'), 11, base=g[t + '/base'])
    assert isinstance(xd, torch.Tensor) and relerr(xd.cpu().numpy(), g[t + '/sample']) < TOL64
    # objectives
    model = vb.GaussianTarget(g[t + '/target_mean'], g[t + '/target_sd'])
    for oname, obj in (('ekl_path', vb.ExclusiveKL(fam, model, 8, use_path_deriv=True)),
                       ('alpha2', vb.AlphaDivergence(fam, model, 8, 2.0))):
        o = '%s/obj/%s' % (t, oname)
        value, grad = obj(vp, base=g[o + '/base'])
        assert relerr(value, g[o + '/value']) < TOL64, (oname, value, g[o + '/value'])
        assert relerr(grad, g[o + '/grad']) < TOL64, oname
    with pytest.raises(NotImplementedError):
        vb.ExclusiveKL(fam, model, 8)(vp)
    with pytest.raises(ValueError):
        fam.sample(vp[:-1], 3)


def test_nvpflow_fit_gaussian_target(vb):
    """A short RMSProp run of the path-derivative ExclusiveKL moves the flow towards a shifted Gaussian target (the
    reference's statistical test style, tests/test_objectives.py): the KL estimate drops and the sample mean follows."""
    dim = 2
    mean, sd = np.array([1.5, -1.0]), np.array([0.7, 1.3])
    model = vb.GaussianTarget(mean, sd)
    layers = [[dim, 8], [8, dim]]
    fam = vb.NVPFlow(layers, layers, nvp_masks(dim, 2), vb.MFGaussian(dim, seed=5), np.zeros(2 * dim), dim)
    rs = np.random.RandomState(3)
    vp = rs.randn(fam.var_param_dim) / 100
    obj = vb.ExclusiveKL(fam, model, 64, use_path_deriv=True)
    v0 = np.mean([obj(vp)[0] for _ in range(5)])
    opt = vb.RMSProp(0.01)
    opt.progress = False
    res = opt.optimize(600, obj, vp)
    vp1 = res['opt_param']
    v1 = np.mean([obj(vp1)[0] for _ in range(5)])
    assert v1 < v0 - 0.5, (v0, v1)
    est_mean, est_cov = fam.mean_and_cov(vp1)
    np.testing.assert_allclose(est_mean, mean, atol=0.35)
