"""CPU: argument checking of the host-side mirror -- same exception class and message as the reference
(file:line of each message in the reference tree is quoted), and the no-CPU-fallback rule."""
import numpy as np
import pytest
import torch

import viabel_b200 as vb

CASES = [
    # (reference file:line, exception, message fragment, callable)
    ('approximations.py:259', ValueError, 'df must be greater than 2', lambda: vb.MFStudentT(3, df=2)),
    ('approximations.py:327', ValueError, 'df must be greater than 2', lambda: vb.MultivariateT(3, df=2)),
    ('_psis.py:145', ValueError, 'More than one log-weight needed.', lambda: vb.psislw(np.zeros(1))),
    ('_psis.py:143', ValueError, 'Argument `lw` must be 1 or 2 dimensional.', lambda: vb.psislw(np.zeros((2, 2, 2)))),
    ('diagnostics.py:173', ValueError, 'alpha must be greater than 1', lambda: vb.divergence_bound(np.zeros(5), alpha=1.0)),
    ('diagnostics.py:133', ValueError, 'must provides samples if moment_bound_fn not given', lambda: vb.wasserstein_bounds(0.5)),
    ('optimization.py:74', ValueError, '"iterate_avg_prop" must be None or between 0 and 1', lambda: vb.RMSProp(0.1, iterate_avg_prop=1.5)),
    ('optimization.py:506', ValueError, 'sgo must be a subclass of StochasticGradientOptimizer', lambda: vb.FASO(object())),
    ('optimization.py:513', ValueError, '"mcse_threshold" must be greater than zero', lambda: vb.FASO(vb.RMSProp(0.1), mcse_threshold=0)),
    ('optimization.py:515', ValueError, '"W_min" must be greater than zero', lambda: vb.FASO(vb.RMSProp(0.1), W_min=0)),
    ('optimization.py:517', ValueError, '"k_check" must be greater than zero', lambda: vb.FASO(vb.RMSProp(0.1), k_check=0)),
    ('optimization.py:519', ValueError, '"ESS_min" must be greater than zero', lambda: vb.FASO(vb.RMSProp(0.1), ESS_min=0)),
    ('optimization.py:675', ValueError, '"rho" must be between zero and one', lambda: vb.RAABBVI(vb.RMSProp(0.1), rho=1.5)),
    ('objectives.py:147', ValueError, "Name of approximation must be one of 'full', 'mean_only'",
     lambda: vb.ExclusiveKL(vb.MFGaussian(2), vb.Model(lambda x: x), 10, hessian_approx_method='bogus')),
    ('convenience.py:64', ValueError, 'either log_density or fit must be specified if objective not given', lambda: vb.bbvi(2)),
    ('convenience.py:71', ValueError, 'if objective is specified, cannot specify fit, log_density, or approx',
     lambda: vb.bbvi(2, log_density=lambda x: x, objective=object())),
    ('convenience.py:124', ValueError, 'either objective or both model and approx must be specified',
     lambda: vb.vi_diagnostics(np.zeros(4))),
]


@pytest.mark.parametrize('where,exc,msg,fn', CASES, ids=[c[0] for c in CASES])
def test_reference_error_behaviour(where, exc, msg, fn):
    with pytest.raises(exc) as e:
        fn()
    assert msg in str(e.value)


@pytest.mark.skipif(torch.cuda.is_available(), reason='only meaningful on a box without a GPU')
def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA: no silent numpy / oracle route."""
    fam = vb.MFGaussian(3)
    with pytest.raises(RuntimeError, match='CUDA'):
        fam.sample(fam.init_param(), 4)
    with pytest.raises(RuntimeError, match='CUDA'):
        vb.psislw(np.random.RandomState(0).randn(100))
    with pytest.raises(RuntimeError, match='CUDA'):
        vb.bbvi(2, log_density=lambda x: -0.5 * (x ** 2).sum(-1), n_iters=5)


# ---- closed forms of the mean-field families run on the host (numpy): pinned against the reference here as well
from conftest import relerr  # noqa: E402


@pytest.mark.parametrize('tag', ['mfg_dfNone_d1', 'mfg_dfNone_d3', 'mfg_dfNone_d8', 'mft_df20_d3', 'mft_df20_d8',
                                 'mft_df5.5_d1', 'mft_df5.5_d3', 'mft_df5.5_d8'])
def test_meanfield_closed_forms_golden(golden, tag):
    """entropy, kl, mean_and_cov, pth_moment, init_param of MFGaussian / MFStudentT
    (approximations.py:207-304) against the unmodified reference -- host code, no device needed."""
    g = golden('families')
    kind, df, d = tag.split('_')
    d = int(d[1:])
    fam = vb.MFGaussian(d) if kind == 'mfg' else vb.MFStudentT(d, float(df[2:]))
    vp, vp1 = g[tag + '/var_param'], g[tag + '/var_param1']
    assert relerr(fam.entropy(vp), g[tag + '/entropy']) < 1e-10
    assert relerr(fam.init_param(), g[tag + '/init_param']) < 1e-15
    if fam.supports_kl:
        assert relerr(fam.kl(vp, vp1), g[tag + '/kl']) < 1e-10
    else:
        with pytest.raises(NotImplementedError):
            fam.kl(vp, vp1)
    mean, cov = fam.mean_and_cov(vp)
    assert relerr(mean, g[tag + '/mean']) < 1e-10 and relerr(cov, g[tag + '/cov']) < 1e-10
    for p in (2, 4):
        key = tag + '/moment%d' % p
        if key in g:
            assert relerr(fam.pth_moment(vp, p), g[key]) < 1e-10
        else:
            with pytest.raises(ValueError):
                fam.pth_moment(vp, p)


def test_host_optimizer_directions_golden(golden):
    """RMSProp / Adam descent directions on numpy gradients (the host path a numpy caller gets;
    optimization.py:188-197, :308-326 incl. Adam's first-step aliasing) against the reference."""
    g = golden('optimizers')
    grads = g['grads']
    for name, mk in (('rmsprop', lambda: vb.RMSProp(0.01)), ('adam', lambda: vb.Adam(0.01)),
                     ('rmsprop_b', lambda: vb.RMSProp(0.01, beta=0.5, jitter=1e-6)),
                     ('adam_b', lambda: vb.Adam(0.01, beta1=0.7, beta2=0.9, jitter=1e-6))):
        opt = mk()
        for i, gr in enumerate(grads):
            assert relerr(opt.descent_direction(gr.copy()), g[name + '/dirs'][i]) < 1e-13, (name, i)


def test_exp_table_constants_accuracy():
    """The table-driven exp of the PSIS passes (csrc/psis.cu exp_nonpos): emulate its arithmetic in numpy with the
    constants parsed from the source.  Guards the reduction constants (256/ln2, the two-part ln2/256) and the
    degree-4 polynomial (and the 7-operation variant of the streaming sums): a wrong digit shows up as an error far above the ~2 ulp of the float64 emulation."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'viabel_b200', 'csrc', 'psis.cu')).read()
    ntab = int(re.search(r'constexpr int kExpTab = (\d+);', src).group(1))
    body = re.search(r'__constant__ double c_expk\[8\] = \{(.*?)\};', src, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    c = [float(t) for t in body.replace('\n', ' ').split(',') if t.strip()]
    assert len(c) == 8 and abs(c[0] - ntab / np.log(2.0)) < 1e-12 * c[0]
    assert abs((c[2] + c[3]) + np.log(2.0) / ntab) < 1e-18              # -(ln2_hi + ln2_lo)/ntab
    shift_bits = int(np.log2(ntab))
    assert re.search(r'const int k = n >> %d;' % shift_bits, src)
    tab = np.exp2(np.arange(ntab) / ntab)
    rs = np.random.RandomState(0)
    x = -np.concatenate([rs.uniform(0, 50, 200000), rs.uniform(0, 700, 50000), 10.0 ** rs.uniform(-12, 0, 50000)])
    n = np.rint(x * c[0])                                  # the 1.5*2^52 shift rounds to nearest
    r = (n * c[2] + x) + n * c[3]
    p = c[4]
    p = p * r + c[5]
    p = p * r + c[6]
    p = p * r + 1.0
    p = p * r + 1.0
    ni = n.astype(np.int64)
    val = np.ldexp(p * tab[ni & (ntab - 1)], (ni >> shift_bits).astype(np.int64))
    rel = np.abs(val - np.exp(x)) / np.exp(x)
    assert rel.max() < 1e-15, rel.max()
    # exp_stream (the two streaming sums): one-step reduction with the correctly rounded -ln2/ntab, degree 3
    assert c[7] == -np.log(2.0) / ntab
    r = n * c[7] + x
    p = c[5]
    p = p * r + c[6]
    p = p * r + 1.0
    p = p * r + 1.0
    val = np.ldexp(p * tab[ni & (ntab - 1)], (ni >> shift_bits).astype(np.int64))
    rel = (val - np.exp(x)) / np.exp(x)
    assert np.abs(rel).max() < 3e-13 and abs(rel.mean()) < 5e-14, (np.abs(rel).max(), rel.mean())


def test_fast_path_numerics_scheme_emulated():
    """The operand scheme of the tensor-core sweep, emulated in numpy: fp16 hi+lo splits of y*X and Theta, three
    products (Xh.Th + Xl.Th + Xh.Tl) accumulated in fp32, link epilogue in fp32, R rounded to fp16, second
    contraction with fp16-exact draws.  Against the float64 oracle it must sit well inside the 1e-4 tolerance,
    while a single bf16 pass (8-bit mantissa) does not -- the reason for the split."""
    rs = np.random.RandomState(3)
    N, d, S = 4096, 48, 32
    X = rs.randn(N, d)
    beta = rs.randn(d) / np.sqrt(d)
    y = np.where(rs.rand(N) < 1 / (1 + np.exp(-X @ beta)), 1.0, -1.0)
    base = rs.randn(S, d).astype(np.float16).astype(np.float64)          # fp16-exact draws
    theta = beta + np.exp(-2.0) * base

    def split16(a):
        hi = a.astype(np.float16)
        lo = (a - hi.astype(np.float64)).astype(np.float16)
        return hi.astype(np.float32), lo.astype(np.float32)

    def sweep(z32):
        z = z32.astype(np.float32)
        t = np.exp2(-np.abs(z) * np.float32(1.4426950408889634))
        sp = np.maximum(-z, 0) + np.log2(1 + t) * np.float32(0.6931471805599453)
        r = (np.where(z >= 0, t, np.float32(1.0)) / (1 + t)).astype(np.float32)
        ll = -sp.astype(np.float64).sum(axis=0)
        r16 = r.astype(np.float16).astype(np.float32)
        Xy = (X * y[:, None])
        T = r16 @ base.astype(np.float32)                                  # [N, d], fp32 accumulate
        ge = (Xy * T.astype(np.float64)).sum(axis=0)
        gmu = Xy.T @ r.astype(np.float64).sum(axis=1)
        return ll, gmu, ge

    Xh, Xl = split16(X * y[:, None])
    Th, Tl = split16(theta)
    z3 = Xh @ Th.T + Xl @ Th.T + Xh @ Tl.T
    zb = (X * y[:, None]).astype(np.float32)
    # bf16 emulation: keep 8 mantissa bits
    def bf16(a):
        b = a.astype(np.float32).view(np.uint32)
        b = ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
        return b.view(np.float32)
    z1 = bf16(zb) @ bf16(theta.astype(np.float32)).T

    a = (X @ theta.T) * y[:, None]
    ll0 = -np.logaddexp(0, -a).sum(axis=0)
    R0 = 1 / (1 + np.exp(a))
    gmu0 = (X * y[:, None]).T @ R0.sum(axis=1)
    ge0 = ((X * y[:, None]) * (R0 @ base)).sum(axis=0)

    def err(z):
        ll, gmu, ge = sweep(z)
        return max(relerr(ll, ll0), relerr(gmu, gmu0), relerr(ge, ge0))

    e3, e1 = err(z3), err(z1)
    assert e3 < 3e-5, e3
    assert e1 > 1e-4, e1
    # the shipped form of GEMM1: the two correction products on e5m2 operands with reciprocal power-of-two scales
    # (X_l 2^ax)(Theta_h 2^-ax) + (X_h 2^-bx)(Theta_l 2^bx), accumulated in the same fp32 sum; E2 takes X as
    # X_h + e5m2(X_l 2^ax) 2^-ax.  Must stay well inside 1e-4, while dropping a correction product does not.
    import torch

    def e5m2(v, scale):
        t = torch.from_numpy(np.ascontiguousarray(v * np.float32(scale), dtype=np.float32))
        return (t.to(torch.float8_e5m2).to(torch.float32).numpy() / np.float32(scale)).astype(np.float32)

    ax, bx = 2, 8                                   # what vb_glm_fast_create derives for max|y X| in [4, 8)
    Xl8 = e5m2(Xl, 2.0 ** ax)
    z8 = Xh @ Th.T + Xl8 @ e5m2(Th, 2.0 ** -ax).T + e5m2(Xh, 2.0 ** -bx) @ e5m2(Tl, 2.0 ** bx).T
    Xe2 = Xh.astype(np.float64) + Xl8.astype(np.float64)
    z = z8.astype(np.float32)
    t = np.exp2(-np.abs(z) * np.float32(1.4426950408889634))
    r = (np.where(z >= 0, t, np.float32(1.0)) / (1 + t)).astype(np.float32)
    sp = np.maximum(-z, 0) + np.log2(1 + t) * np.float32(0.6931471805599453)
    T = r.astype(np.float16).astype(np.float32) @ base.astype(np.float32)
    e8 = max(relerr(-sp.astype(np.float64).sum(axis=0), ll0), relerr(Xe2.T @ r.astype(np.float64).sum(axis=1), gmu0),
             relerr((Xe2 * T.astype(np.float64)).sum(axis=0), ge0))
    e2 = err(Xh @ Th.T + Xh @ Tl.T)                 # without X_l . Theta_h
    assert e8 < 4e-5, e8
    assert e2 > 2 * e8, (e2, e8)


def test_product_never_touches_the_oracle_and_needs_its_library(tmp_path):
    """The oracle is test infrastructure: no product module may mention it; and without the built CUDA library the
    package must refuse to import (no silent fallback)."""
    import glob
    import os
    import shutil
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in glob.glob(os.path.join(root, 'viabel_b200', '*.py')) + glob.glob(os.path.join(root, 'viabel_b200', 'csrc', '*')):
        if os.path.isfile(f):
            assert 'oracle' not in open(f, errors='ignore').read(), f
    pkg = tmp_path / 'viabel_b200'
    pkg.mkdir()
    for f in glob.glob(os.path.join(root, 'viabel_b200', '*.py')):
        shutil.copy(f, pkg)                                  # the Python side only: no libviabel_b200.so
    out = subprocess.run([sys.executable, '-c', 'import viabel_b200'], cwd=str(tmp_path), capture_output=True, text=True,
                         timeout=120)
    assert out.returncode != 0 and 'ImportError' in out.stderr and 'libviabel_b200' in out.stderr


def test_stan_model_adapter_host_side():
    """StanModel (models.py:80-100) wraps a PyStan fit object: construction and `constrain` are host-only (the
    log density itself needs CUDA tensors: tests/test_gpu_elbo.py::test_stan_model_adapter)."""
    import viabel_b200 as vb

    class Fit:                                   # the three methods of StanFit4model the reference uses
        def log_prob(self, x):
            return -0.5 * float(np.sum(np.asarray(x) ** 2))

        def grad_log_prob(self, x):
            return -np.asarray(x)

        def constrain_pars(self, x):
            return {'theta': np.asarray(x)}

    m = vb.StanModel(Fit())
    assert isinstance(m, vb.Model) and not m.supports_tempering
    assert np.array_equal(m.constrain(np.arange(3.0))['theta'], np.arange(3.0))
    with pytest.raises(NotImplementedError):
        m.set_inverse_temperature(0.5)
