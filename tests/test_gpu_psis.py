"""GPU parity for PSIS / divergence bounds / vi_diagnostics against the goldens produced by the
reference's own _psis.py and diagnostics.py (run unmodified) and against the numpy oracle.
Tolerances: k-hat, bounds 1e-10 relative; tail index set bit-exact."""
import contextlib
import io

import numpy as np
import pytest
import torch

from conftest import relerr
from _problems import PSIS_CASES, diag_problem, psis_case

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope='module')
def vb():
    import viabel_b200
    return viabel_b200


@pytest.fixture(scope='module')
def vo():
    from oracle import viabel_oracle
    return viabel_oracle


def _check_case(vb, g, name, out, k, tail=None):
    kref = g[name + '/khat']
    assert np.array_equal(np.isinf(k), np.isinf(kref))
    fin = np.isfinite(kref)
    assert relerr(np.asarray(k)[fin], kref[fin]) < TOL
    stride = int(g[name + '/out_stride'])
    if name != 'ties_3e4':
        np.testing.assert_allclose(out[::stride], g[name + '/out_sub'], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(np.sort(out, axis=0)[::stride], g[name + '/out_sorted_sub'], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(np.max(out, axis=0), g[name + '/out_max'], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(np.min(out, axis=0), g[name + '/out_min'], rtol=1e-10, atol=1e-10)
    if tail is not None:
        assert np.array_equal(tail, g[name + '/tail_idx'])          # bit-exact index set


@pytest.mark.parametrize('name', PSIS_CASES)
def test_psislw_golden(vb, golden, name):
    g = golden('psis')
    lw = psis_case(name)
    if lw.ndim == 1:
        out, k, tail_idx, tail_rank = vb.psislw(lw, return_tail=True)
        assert isinstance(out, np.ndarray) and out is not lw
        _check_case(vb, g, name, out, k, tail_idx)
        # rank order: tail values sorted by (value, index)
        x = lw - lw.max()
        order = np.lexsort((tail_idx, x[tail_idx]))
        assert np.array_equal(np.argsort(order, kind='stable'), tail_rank)
        if name + '/d2' in g:
            d2, elbo = vb.divergence_bound(out, return_log_norm_bound=True)
            assert relerr(d2, g[name + '/d2']) < 1e-9 and relerr(elbo, g[name + '/elbo']) < TOL
            for alpha in (1.5, 3.0):
                assert relerr(vb.divergence_bound(out, alpha=alpha), g[name + '/dalpha%.1f' % alpha]) < 1e-9
    else:
        out, k = vb.psislw(lw)
        assert out.flags.f_contiguous and k.shape == (lw.shape[1],)
        _check_case(vb, g, name, out, k)
        np.testing.assert_allclose(np.exp(out).sum(axis=0), 1.0, rtol=1e-12)


def test_psislw_modes(vb, golden):
    g = golden('psis')
    name = 't5_t7_1e5'
    lw = psis_case(name)
    # device tensor in -> device tensor out, in place
    t = torch.as_tensor(lw, device='cuda')
    out, k = vb.psislw(t, overwrite_lw=True)
    assert out.data_ptr() == t.data_ptr()
    _check_case(vb, g, name, out.cpu().numpy(), k)
    # exact (full radix-select) mode gives the same answer as the sampled-threshold mode
    t = torch.as_tensor(lw, device='cuda')
    o2 = torch.empty_like(t)
    _, res, ti, tr = vb.psislw_device(t, o2, want_tail=True, exact=True)
    res = res.cpu().numpy()
    assert res[6] == 0
    _check_case(vb, g, name, o2.cpu().numpy(), res[0], ti[:int(res[2])].cpu().numpy())
    # k-hat only (no output array), and the fused CUBO / ELBO moments
    _, res2, _, _ = vb.psislw_device(t, None)
    assert relerr(res2.cpu().numpy()[0], g[name + '/khat']) < TOL
    n = lw.size
    cubo = np.log(res[8] / n) / 2 - res[4]
    elbo = res[7] / n - res[4]
    assert relerr(2 * (cubo - elbo), g[name + '/d2']) < 1e-9
    assert relerr(elbo, g[name + '/elbo']) < TOL
    # numpy in-place
    a = lw.copy()
    out, _ = vb.psislw(a, overwrite_lw=True)
    assert out is a
    with pytest.raises(ValueError):
        vb.psislw(np.zeros(1))
    with pytest.raises(ValueError):
        vb.psislw(np.zeros((2, 2, 2)))


@pytest.mark.parametrize('name', ['t5_t7_1e5', 't3_t30_2e5', 't50_t5_1e5', 'tiny_20', 'small_100', 'ties_3e4',
                                  'underflow_5000', 'all_equal_50', 'big_1e6'])
@pytest.mark.parametrize('variant', ['plain', 'exact', 'unaligned'])
def test_psislw_moments_only_one_pass(vb, vo, name, variant):
    """k-hat / bounds only (no output array) takes ONE pass over the draws: its k-hat, cutoff, log-sum-exp and the
    CUBO / ELBO sums must equal the two-pass results and the oracle's (moments of v = out + lse) to 1e-10."""
    lw = psis_case(name)
    if variant == 'unaligned':
        buf = torch.zeros(lw.size + 1, device='cuda', dtype=torch.float64)
        buf[1:] = torch.as_tensor(lw, device='cuda')
        t = buf[1:]                                        # 8-byte aligned only: the scalar loop
    else:
        t = torch.as_tensor(lw, device='cuda')
    exact = variant == 'exact'
    out = torch.empty(lw.size, device='cuda', dtype=torch.float64)
    _, res2, _, _ = vb.psislw_device(t, out, exact=exact)
    _, res1, _, _ = vb.psislw_device(t, None, exact=exact)
    r1, r2 = res1.cpu().numpy(), res2.cpu().numpy()
    assert r1[6] == r2[6]
    if r2[6] != 0:                                         # sampled threshold missed (tiny / tied inputs): rerun exact
        _, res2, _, _ = vb.psislw_device(t, out, exact=True)
        _, res1, _, _ = vb.psislw_device(t, None, exact=True)
        r1, r2 = res1.cpu().numpy(), res2.cpu().numpy()
    assert r1[6] == 0 and r2[6] == 0
    for slot in (0, 1, 2, 3, 4, 5, 9, 11):                 # khat, sigma, n2, cutoff, lse, max, M, smoothed
        assert (r1[slot] == r2[slot]) or relerr(r1[slot], r2[slot]) < 1e-12, (slot, r1[slot], r2[slot])
    with np.errstate(all='ignore'):
        o_ref, k_ref = vo.psislw_1d(lw)[:2]
    v = o_ref + r2[4]
    sumv, sume = v.sum(), np.exp(2.0 * v).sum()
    for got in (r1, r2):
        assert relerr(got[7], sumv) < TOL and relerr(got[8], sume) < TOL, (got[7], sumv, got[8], sume)
    if np.isfinite(k_ref):
        assert relerr(r1[0], k_ref) < TOL


def test_psisloo_golden(vb, golden):
    """PSIS leave-one-out (_psis.py:69-110) against the unmodified reference: every column of -log_lik is one
    PSIS problem; loo, the per-term loos and the tail indices to 1e-10."""
    g = golden('psisloo')
    ll = g['log_lik']
    keep = ll.copy()
    loo, loos, ks = vb.psisloo(ll)
    assert np.array_equal(ll, keep)                        # the caller's array is not touched
    assert relerr(loo, g['loo']) < TOL and relerr(loos, g['loos']) < TOL and relerr(ks, g['ks']) < TOL
    with pytest.raises(ValueError):
        vb.psisloo(np.zeros(5))


@pytest.mark.parametrize('n', [5000, 300001])
def test_psislw_nan_and_minus_inf(vb, vo, n):
    """Non-finite log-weights, as the reference treats them (_psis.py run unmodified: a NaN makes every output NaN
    and k-hat +inf -- the tail {x > cutoff} is empty; -inf weights are ordinary draws of weight zero)."""
    rs = np.random.RandomState(n)
    base = rs.standard_t(4, size=n)
    for pos in (1234, n - 1):                              # vectorised main loop / scalar remainder of pass A
        for nan in (np.nan, -np.nan, np.inf - np.inf):
            lw = base.copy()
            lw[pos] = nan
            for out_wanted in (True, False):
                t = torch.as_tensor(lw, device='cuda')
                out = torch.empty_like(t) if out_wanted else None
                for exact in (False, True):                # status 1 = sampled threshold missed: the API reruns exact
                    _, res, _, _ = vb.psislw_device(t, out, exact=exact)
                    r = res.cpu().numpy()
                    if r[6] != 1:
                        break
                assert r[6] == 0 and np.isposinf(r[0]), (pos, r[0], r[6])
                if out_wanted:
                    assert bool(torch.isnan(out).all())
            out, k = vb.psislw(lw)
            assert np.isposinf(k) and np.isnan(out).all()
    lw = base.copy()
    lw[[7, 1234, n - 1]] = -np.inf
    out, k = vb.psislw(lw)
    with np.errstate(all='ignore'):
        o_ref, k_ref = vo.psislw_1d(lw)[:2]
    assert relerr(k, k_ref) < TOL
    fin = np.isfinite(o_ref)
    assert np.array_equal(np.isneginf(out), np.isneginf(o_ref)) and fin.sum() == n - 3
    np.testing.assert_allclose(out[fin], o_ref[fin], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize('n,dfp,dfq', [(3000000, 4, 9), (10000000, 3, 30)])
def test_psislw_large_vs_oracle(vb, vo, n, dfp, dfq):
    g = torch.Generator(device='cuda')
    g.manual_seed(n)
    z = torch.randn(n, generator=g, device='cuda', dtype=torch.float64)
    c = torch.distributions.Chi2(torch.tensor(float(dfq), device='cuda', dtype=torch.float64)).sample((n,))
    s = z / torch.sqrt(c / dfq)
    lw = (-(dfp + 1) / 2 * torch.log1p(s * s / dfp) + (dfq + 1) / 2 * torch.log1p(s * s / dfq)).contiguous()
    out = torch.empty_like(lw)
    _, res, ti, _ = vb.psislw_device(lw, out, want_tail=True)
    res = res.cpu().numpy()
    assert res[6] == 0
    with np.errstate(all='ignore'):
        o_ref, k_ref, tail_ref, _ = vo.psislw_1d(lw.cpu().numpy(), return_tail=True)
    assert relerr(res[0], k_ref) < TOL
    assert np.array_equal(ti[:int(res[2])].cpu().numpy(), tail_ref)
    np.testing.assert_allclose(out.cpu().numpy(), o_ref, rtol=1e-10, atol=1e-10)


def test_psislw_1e8_vs_oracle(vb, vo):
    """BASELINE configs[4] at FULL size (n = 1e8 draws) against the oracle, not only by properties: k-hat to
    1e-10, the tail index set bit-exact, the smoothed normalised weights to 1e-10, and the divergence bound."""
    n = 100000000
    g = torch.Generator(device='cuda')
    g.manual_seed(20260119)
    lw = torch.zeros(n, device='cuda', dtype=torch.float64)
    for _ in range(2):                                  # t_10 target under a t_40 proposal, two coordinates
        z = torch.randn(n, generator=g, device='cuda', dtype=torch.float64)
        lw += -5.5 * torch.log1p(z * z / 10.0) + 20.5 * torch.log1p(z * z / 40.0)
        del z
    out = torch.empty_like(lw)
    _, res, ti, _ = vb.psislw_device(lw, out, want_tail=True)
    res = res.cpu().numpy()
    assert res[6] == 0 and res[9] == 30000
    host = lw.cpu().numpy()
    del lw
    with np.errstate(all='ignore'):
        o_ref, k_ref, tail_ref, _ = vo.psislw_1d(host, return_tail=True)
    del host
    assert relerr(res[0], k_ref) < TOL
    assert np.array_equal(ti[:int(res[2])].cpu().numpy(), tail_ref)
    d2 = vb.divergence_bound(out)
    o = out.cpu().numpy()
    del out
    assert np.max(np.abs(o - o_ref)) < 1e-10 * np.max(np.abs(o_ref))
    assert relerr(d2, vo.divergence_bound(o_ref)[0]) < TOL


def test_psislw_1e8_properties(vb):
    """BASELINE configs[4] size: size-independent properties (normalisation, tail size, clamp,
    idempotent k-hat under a shift of the weights)."""
    n = 100000000
    g = torch.Generator(device='cuda')
    g.manual_seed(5)
    z = torch.randn(n, generator=g, device='cuda', dtype=torch.float64)
    lw = -0.5 * z * z * 0.35 + 0.1 * z          # light right tail
    del z
    out = torch.empty_like(lw)
    _, res, ti, tr = vb.psislw_device(lw, out, want_tail=True)
    r = res.cpu().numpy()
    assert r[6] == 0 and r[9] == 30000 and 0 < r[2] <= 30000
    assert abs(float(torch.logsumexp(out, 0))) < 1e-9
    assert float(out.max()) <= -r[4] + 1e-12
    idx = ti[:int(r[2])]
    assert bool((idx[1:] > idx[:-1]).all())
    shifted = lw + 123.456
    _, res2, _, _ = vb.psislw_device(shifted, None)
    assert relerr(res2.cpu().numpy()[0], r[0]) < 1e-4        # the shift re-rounds every weight


def test_diagnostics_golden(vb, golden):
    g = golden('diagnostics')
    samples, lw = diag_problem()
    keys = ['W1', 'W2', 'mean_error', 'std_error', 'cov_error', 'd2', 'log_norm_bound']
    for alpha in (1.5, 2.0, 3.0):
        assert relerr(vb.divergence_bound(lw, alpha=alpha), g['dalpha%.1f' % alpha]) < TOL
        assert relerr(vb.divergence_bound(lw, alpha=alpha, log_norm_bound=0.0), g['dalpha%.1f_lnb0' % alpha]) < TOL
    wb = vb.wasserstein_bounds(0.7, samples=samples)
    assert relerr([wb['W1'], wb['W2']], g['wb_samples']) < TOL
    wb = vb.wasserstein_bounds(0.7, samples=samples[:, 0])
    assert relerr([wb['W1'], wb['W2']], g['wb_samples_1d']) < TOL
    wb = vb.wasserstein_bounds(0.7, moment_bound_fn=lambda p: 3.0 * p)
    assert relerr([wb['W1'], wb['W2']], g['wb_fn']) < TOL
    res = vb.all_diagnostics(lw, samples=samples)
    assert relerr([res[k] for k in keys], g['all_samples']) < TOL
    res = vb.all_diagnostics(torch.as_tensor(lw, device='cuda'), moment_bound_fn=lambda p: 2.5 * p, q_var=1.7)
    assert relerr([res[k] for k in keys], g['all_fn_scalar']) < TOL
    res = vb.all_diagnostics(lw, samples=samples, q_var=np.cov(samples.T) * 1.1, p_var=0.9, log_norm_bound=-1.5)
    assert relerr([res[k] for k in keys], g['all_full']) < TOL
    eb = vb.error_bounds(W1=0.3, W2=0.5, q_var=2.0)
    assert relerr([eb['mean_error'], eb['std_error'], eb['cov_error']], g['error_bounds']) < TOL
    with pytest.raises(ValueError):
        vb.divergence_bound(lw, alpha=1.0)
    with pytest.raises(ValueError):
        vb.wasserstein_bounds(0.5)


def test_vi_diagnostics_golden(vb, golden):
    """convenience.py:97-179 end to end with the reference's own draws injected."""
    from viabel_b200.convenience import _vi_diagnostics
    g = golden('diagnostics')
    eps = g['vi_eps']
    for name in ('matched', 'narrow', 'wide'):
        vp = g['vi_%s/var_param' % name]
        mean = torch.as_tensor(g['vi_%s/target_mean' % name], device='cuda')
        sd = torch.as_tensor(g['vi_%s/target_sd' % name], device='cuda')

        def log_p(x):
            z = (x - mean) / sd
            return (-0.5 * z * z - torch.log(sd) - 0.5 * np.log(2 * np.pi)).sum(dim=1)
        with contextlib.redirect_stdout(io.StringIO()):
            res = _vi_diagnostics(vp, vb.Model(log_p), vb.MFGaussian(4), eps.shape[0], base=eps)
        assert relerr(res['khat'], g['vi_%s/khat' % name]) < 1e-9
        assert res['samples'].shape == (4, eps.shape[0])
        np.testing.assert_allclose(res['smoothed_log_weights'][::20], g['vi_%s/slw_sub' % name], rtol=1e-9, atol=1e-9)
        for key in ('W1', 'W2', 'mean_error', 'std_error', 'cov_error', 'd2', 'log_norm_bound'):
            if 'vi_%s/%s' % (name, key) in g:
                assert relerr(res[key], g['vi_%s/%s' % (name, key)]) < 1e-8, (name, key)
            else:
                assert key not in res


def test_bbvi_and_vi_diagnostics_scenarios(vb):
    """tests/test_convenience.py scenarios: bbvi (RAABBVI / FASO / RMSProp) recovers a Gaussian,
    argument validation, and the three k-hat regimes of vi_diagnostics."""
    mean = torch.tensor([3., -4.], dtype=torch.float64, device='cuda')
    sd = torch.tensor([2., 5.], dtype=torch.float64, device='cuda')

    def log_p(x):
        z = (x - mean) / sd
        return (-0.5 * z * z - torch.log(sd) - 0.5 * np.log(2 * np.pi)).sum(dim=1)

    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        for adaptive, fixed_lr, S in ((True, True, 1000), (True, False, 1000), (False, True, 50)):
            np.random.seed(851)
            approx = vb.MFGaussian(2)
            objective = vb.ExclusiveKL(approx, vb.Model(log_p), S)
            res = vb.bbvi(2, objective=objective, adaptive=adaptive, fixed_lr=fixed_lr, n_iters=15000,
                          RAABBVI_kwargs=dict(mcse_threshold=.005, accuracy_threshold=.005),
                          FASO_kwargs=dict(mcse_threshold=.005), RMS_kwargs=dict())
            est_mean, est_cov = res['objective'].approx.mean_and_cov(res['opt_param'])
            dec = 2 if adaptive else 1
            np.testing.assert_almost_equal(est_mean, [3., -4.], decimal=dec)
            np.testing.assert_almost_equal(np.sqrt(np.diag(est_cov)), [2., 5.], decimal=dec)
    with pytest.raises(ValueError):
        vb.bbvi(2)
    with pytest.raises(ValueError):
        vb.bbvi(2, objective=True, fit=True)
    with pytest.raises(ValueError):
        vb.bbvi(2, log_density=True, fit=True)
    with pytest.raises(ValueError):
        vb.bbvi(2, objective=True, log_density=True)

    def std_normal(scale):
        def f(x):
            return (-0.5 * (x / scale) ** 2 - np.log(scale) - 0.5 * np.log(2 * np.pi)).sum(dim=1)
        return f
    with contextlib.redirect_stdout(out):
        approx = vb.MFGaussian(2)
        vp = np.zeros(4)                      # q = N(0, I)
        d1 = vb.vi_diagnostics(vp, model=vb.Model(std_normal(1.0)), approx=approx)
        d2 = vb.vi_diagnostics(vp, model=vb.Model(std_normal(3.0)), approx=approx)
        d3 = vb.vi_diagnostics(vp, model=vb.Model(std_normal(0.5)), approx=approx)
    assert d1['khat'] < .1 and d1['d2'] < 0.1
    assert d2['khat'] > 0.7 and 'd2' not in d2
    assert d3['khat'] < 0 and d3['d2'] > 2
    with pytest.raises(ValueError):
        vb.vi_diagnostics(vp, approx=approx)
    with pytest.raises(ValueError):
        vb.vi_diagnostics(vp, model=vb.Model(std_normal(1.0)), approx=approx, n_samples=0)


# ---- draw-sharded PSIS: shards driven in one process, exchange done by hand ------------
def _sharded(vb, lw, cuts, exact=False):
    from viabel_b200._psis import PsisShard
    x = torch.as_tensor(lw, dtype=torch.float64, device='cuda')
    n, world = x.numel(), len(cuts) - 1
    shards = [PsisShard(x[cuts[r]:cuts[r + 1]].contiguous(), cuts[r], n, 1.0, world, r) for r in range(world)]
    recs = torch.cat([s.local(exact) for s in shards])           # what the all-gather would deliver
    outs, res, mom = [], None, np.zeros(2)
    for s in shards:
        s.global_(recs)
        o = torch.empty_like(s.lw)
        rr = s.apply(o).cpu().numpy()
        assert rr[6] == 0
        mom += rr[7:9]
        outs.append(o)
        if res is None:
            res = rr
        else:
            assert np.array_equal(rr[:7], res[:7]) and np.array_equal(rr[9:12], res[9:12])    # replicated stage
    return torch.cat(outs).cpu().numpy(), res, mom


@pytest.mark.parametrize('name,cuts', [('t5_t7_1e5', [0, 50000, 100000]), ('t5_t7_1e5', [0, 1000, 61234, 100000]),
                                       ('t3_t30_2e5', [0, 66667, 200000]), ('ties_3e4', [0, 15000, 30000]),
                                       ('small_100', [0, 40, 100]), ('small_100', [0, 10, 22, 100]),
                                       ('underflow_5000', [0, 2500, 5000]), ('all_equal_50', [0, 20, 50]),
                                       ('big_1e6', [0, 250000, 500000, 750000, 1000000])])
@pytest.mark.parametrize('exact', [False, True])
def test_psislw_sharded_matches_single(vb, name, cuts, exact):
    lw = psis_case(name)
    from viabel_b200._psis import psislw_device
    x = torch.as_tensor(lw, dtype=torch.float64, device='cuda')
    o1 = torch.empty_like(x)
    _, r1, _, _ = psislw_device(x, o1, 1.0, exact=exact)
    r1 = r1.cpu().numpy()
    assert r1[6] == 0
    out, res, mom = _sharded(vb, lw, cuts, exact)
    assert res[0] == r1[0] and res[1] == r1[1] and res[2] == r1[2] and res[3] == r1[3]      # k-hat, sigma, n2, cutoff
    assert abs(res[4] - r1[4]) < 1e-12 and res[5] == r1[5]
    np.testing.assert_allclose(out, o1.cpu().numpy(), rtol=0, atol=1e-11)
    np.testing.assert_allclose(mom, r1[7:9], rtol=1e-11)


def test_psislw_exact_mode_all_equal(vb):
    """Exact mode, no value strictly above the cutoff: the maximum is the cutoff itself."""
    from viabel_b200._psis import psislw_device
    x = torch.full((50,), -3.25, dtype=torch.float64, device='cuda')
    o = torch.empty_like(x)
    _, r, _, _ = psislw_device(x, o, 1.0, exact=True)
    r = r.cpu().numpy()
    assert r[6] == 0 and np.isinf(r[0]) and r[5] == -3.25
    np.testing.assert_allclose(o.cpu().numpy(), -np.log(50.0), rtol=1e-14)


def test_psislw_sharded_uneven_tail_owner(vb):
    """All of the global tail lives on one shard; the other shard's draws are all below the cutoff."""
    rng = np.random.default_rng(5)
    lo = rng.standard_normal(40000) - 50.0
    hi = rng.standard_t(3, 60000) * 2.0
    lw = np.concatenate([lo, hi])
    out1, k1 = vb.psislw(lw)
    out, res, _ = _sharded(vb, lw, [0, 40000, 100000])
    assert res[0] == k1
    np.testing.assert_allclose(out, out1, rtol=0, atol=1e-11)


def test_psislw_sharded_api_single_process(vb):
    """psislw_sharded without a process group = psislw."""
    lw = psis_case(PSIS_CASES[0])
    lw = lw if lw.ndim == 1 else lw[:, 0].copy()
    out1, k1 = vb.psislw(lw)
    out, k, res = vb.psislw_sharded(torch.as_tensor(lw, device='cuda'))
    assert k == k1
    np.testing.assert_allclose(out.cpu().numpy(), out1, rtol=0, atol=1e-12)


@pytest.mark.parametrize('target_kind', ['student', 'gauss', 'generic'])
@pytest.mark.parametrize('family', ['mft', 'mfg'])
def test_vi_diagnostics_streaming_matches_materialised(vb, target_kind, family):
    """Streaming vi_diagnostics (draws regenerated by Philox offset, never stored; convenience.py:136-179 is the
    materialising reference flow) gives the same k-hat, smoothed weights and bounds as the path that keeps
    samples[n, d] -- through the fused draw+density kernel for the built-in product targets and through chunks for
    a user model."""
    d, n = 7, 60011
    rs = np.random.RandomState(5)
    loc, scale = rs.randn(d), np.exp(0.2 * rs.randn(d))
    vp = np.concatenate([loc + 0.1 * rs.randn(d), np.log(scale) + 0.05 * rs.randn(d)])
    if target_kind == 'student':
        model = vb.StudentTTarget(loc, scale, 10.0)
    elif target_kind == 'gauss':
        model = vb.GaussianTarget(loc, scale)
    else:
        lt, st = torch.as_tensor(loc, device='cuda'), torch.as_tensor(scale, device='cuda')
        model = vb.Model(lambda x: (-0.5 * ((x - lt) / st) ** 2 - torch.log(st) - 0.5 * np.log(2 * np.pi)).sum(dim=1))
    mk = (lambda: vb.MFStudentT(d, 40, seed=3)) if family == 'mft' else (lambda: vb.MFGaussian(d, seed=3))
    a1, a2 = mk(), mk()
    with contextlib.redirect_stdout(io.StringIO()):
        full = vb.vi_diagnostics(vp, model=model, approx=a1, n_samples=n, keep_samples=True)
        stream = vb.vi_diagnostics(vp, model=model, approx=a2, n_samples=n, keep_samples=False)
    assert stream['samples'] is None and a1._offset == a2._offset
    assert relerr(stream['khat'], full['khat']) < 1e-9
    np.testing.assert_allclose(stream['smoothed_log_weights'].cpu().numpy(), full['smoothed_log_weights'], rtol=1e-9, atol=1e-9)
    for key in ('d2', 'W1', 'W2', 'mean_error', 'std_error', 'cov_error', 'log_norm_bound'):
        assert relerr(stream[key], full[key]) < 1e-8, key
