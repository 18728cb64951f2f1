"""CPU oracle for the viabel hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (viabel_b200/) never does, and it
fails loudly when its CUDA library is missing.

What this is: a float64 numpy restatement of the reference algorithm
(jhuggins/viabel v0.5.2, /root/reference/viabel/...) for the per-iteration
stochastic ELBO / alpha-divergence gradient and the PSIS / divergence-bound
diagnostics.  Each function cites the reference lines it follows.  The reference
obtains gradients from the third-party `autograd` package (requirements.txt:3,
autograd~=1.3, not vendored); here they are written out analytically
(SURVEY.md App. A).

Parity pin: tests/golden/*.npz hold outputs of the UNMODIFIED reference classes
run in the build container (oracle/make_golden.py; autograd/paragami replaced by
the torch-backed stand-ins in oracle/refshim).  tests/test_oracle_golden.py checks
every function here against them.  _psis.py / diagnostics.py of the reference
depend on numpy only and were run as-is.

Random draws are never reproduced: every sampler takes the base draws
(eps / t / (chi2, z)) as an argument ("draw injection").
"""
import math

import numpy as np
from scipy import special as sps

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------
# parameter layouts (paragami patterns; approximations.py:185-189, 315-319)
# --------------------------------------------------------------------------
def mf_unpack(var_param, dim):
    """[mu(d), log_sigma(d)] -- PatternDict insertion order (approximations.py:185-189)."""
    vp = np.asarray(var_param, dtype=np.float64)
    return vp[:dim], vp[dim:2 * dim]


def mvt_param_dim(dim):
    return dim + dim * (dim + 1) // 2


def mvt_unpack(var_param, dim):
    """[mu(d), rowmajor-tril(F)] with L = tril(F,-1)+diag(exp(diag F)), Sigma = L L^T
    (approximations.py:315-319; paragami PSDSymmetricMatrixPattern, diag_lb=0)."""
    vp = np.asarray(var_param, dtype=np.float64)
    mu = vp[:dim]
    F = np.zeros((dim, dim))
    F[np.tril_indices(dim)] = vp[dim:]
    L = np.tril(F, -1) + np.diag(np.exp(np.diag(F)))
    return mu, L


def mvt_pack(mu, Sigma):
    d = len(mu)
    L = np.linalg.cholesky(Sigma)
    F = L.copy()
    F[np.diag_indices(d)] = np.log(np.diag(L))
    return np.concatenate([np.asarray(mu, float), F[np.tril_indices(d)]])


def sym_sqrt(Sigma):
    """Symmetric PSD square root; equals scipy.linalg.sqrtm on PSD input
    (approximations.py:348)."""
    w, V = np.linalg.eigh(Sigma)
    return (V * np.sqrt(np.maximum(w, 0.0))) @ V.T, w, V


# --------------------------------------------------------------------------
# MFGaussian (approximations.py:192-251)
# --------------------------------------------------------------------------
def mfg_init_param(dim):
    """approximations.py:207-210"""
    return np.concatenate([np.zeros(dim), 2.0 * np.ones(dim)])


def mfg_sample(var_param, eps):
    """theta = mu + exp(log_sigma) * eps  (approximations.py:212-216)"""
    mu, ls = mf_unpack(var_param, eps.shape[1])
    return mu + np.exp(ls) * eps


def mfg_entropy(var_param, dim):
    """approximations.py:218-220"""
    _, ls = mf_unpack(var_param, dim)
    return 0.5 * dim * (1.0 + LOG_2PI) + np.sum(ls)


def mfg_kl(vp0, vp1, dim):
    """KL(q0 || q1), approximations.py:222-229"""
    m0, l0 = mf_unpack(vp0, dim)
    m1, l1 = mf_unpack(vp1, dim)
    dl = l0 - l1
    return 0.5 * np.sum(np.exp(2 * dl) + (m0 - m1) ** 2 / np.exp(2 * l1) - 2 * dl - 1.0)


def mfg_log_density(var_param, x):
    """sum_j norm.logpdf (approximations.py:231-236); 1-D x is promoted."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    mu, ls = mf_unpack(var_param, x.shape[1])
    z = (x - mu) / np.exp(ls)
    return np.sum(-0.5 * z * z - ls - 0.5 * LOG_2PI, axis=1)


def mfg_mean_and_cov(var_param, dim):
    """approximations.py:238-240 (dense diagonal matrix)"""
    mu, ls = mf_unpack(var_param, dim)
    return mu.copy(), np.diag(np.exp(2 * ls))


def mfg_pth_moment(var_param, dim, p):
    """approximations.py:242-248"""
    if p not in (2, 4):
        raise ValueError('p = {} is not a supported moment'.format(p))
    _, ls = mf_unpack(var_param, dim)
    v = np.exp(2 * ls)
    return np.sum(v) if p == 2 else 2 * np.sum(v ** 2) + np.sum(v) ** 2


# --------------------------------------------------------------------------
# MFStudentT (approximations.py:254-312)
# --------------------------------------------------------------------------
def mft_sample(var_param, tdraws):
    """approximations.py:270-274"""
    return mfg_sample(var_param, tdraws)


def mft_entropy(var_param, dim):
    """sum(log_sigma), df-only constants dropped (approximations.py:276-279)"""
    return np.sum(mf_unpack(var_param, dim)[1])


def student_t_logpdf(z, df):
    """scipy.stats.t.logpdf for a standardised variate."""
    return (sps.gammaln(0.5 * (df + 1.0)) - sps.gammaln(0.5 * df)
            - 0.5 * math.log(df * math.pi) - 0.5 * (df + 1.0) * np.log1p(z * z / df))


def mft_log_density(var_param, x, df):
    """approximations.py:281-286"""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    mu, ls = mf_unpack(var_param, x.shape[1])
    z = (x - mu) / np.exp(ls)
    return np.sum(student_t_logpdf(z, df) - ls, axis=1)


def mft_mean_and_cov(var_param, dim, df):
    """approximations.py:288-292"""
    mu, ls = mf_unpack(var_param, dim)
    return mu.copy(), df / (df - 2.0) * np.diag(np.exp(2 * ls))


def mft_pth_moment(var_param, dim, df, p):
    """approximations.py:294-307"""
    if p not in (2, 4) or not p < df:
        raise ValueError('p = {} is not a supported moment'.format(p))
    _, ls = mf_unpack(var_param, dim)
    s2 = np.exp(2 * ls)
    c = df / (df - 2.0)
    if p == 2:
        return c * np.sum(s2)
    return c ** 2 * (2 * (df - 1.0) / (df - 4.0) * np.sum(s2 ** 2) + np.sum(s2) ** 2)


# --------------------------------------------------------------------------
# MultivariateT (approximations.py:322-382; _distributions.py:7-38)
# --------------------------------------------------------------------------
def mvt_init_param(dim):
    """mu = 0, Sigma = 10 I (approximations.py:337-340)"""
    return mvt_pack(np.zeros(dim), 10.0 * np.eye(dim))


def mvt_sample(var_param, chi2, z, df):
    """theta = mu + (z @ sqrtm(Sigma)) / sqrt(chi2/df)[:,None]  (approximations.py:342-349).
    The reference draws chi2 BEFORE z."""
    d = z.shape[1]
    mu, L = mvt_unpack(var_param, d)
    A, _, _ = sym_sqrt(L @ L.T)
    u = np.sqrt(chi2 / df)
    return mu + (z @ A) / u[:, None]


def mvt_entropy(var_param, dim):
    """0.5*log(det(Sigma)) via det (approximations.py:351-354) -- overflows for large d,
    exactly like the reference; mvt_entropy_stable is what the product computes."""
    _, L = mvt_unpack(var_param, dim)
    with np.errstate(over='ignore'):
        return 0.5 * np.log(np.linalg.det(L @ L.T))


def mvt_entropy_stable(var_param, dim):
    _, L = mvt_unpack(var_param, dim)
    return np.sum(np.log(np.diag(L)))


def mvt_log_density(var_param, x, df):
    """_distributions.py:7-38: eigh(Sigma), eigenvalues <= 1e-10 get a zero inverse but
    still enter the log pseudo-determinant."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    d = x.shape[1]
    mu, L = mvt_unpack(var_param, d)
    w, V = np.linalg.eigh(L @ L.T)
    winv = np.where(np.abs(w) <= 1e-10, 0.0, 1.0 / np.where(w == 0, 1.0, w))
    U = V * np.sqrt(winv)
    maha = np.sum(((x - mu) @ U) ** 2, axis=-1)
    return (sps.gammaln(0.5 * (df + d)) - sps.gammaln(0.5 * df) - 0.5 * d * math.log(math.pi * df)
            - 0.5 * np.sum(np.log(w)) - 0.5 * (df + d) * np.log(1.0 + maha / df))


def mvt_mean_and_cov(var_param, dim, df):
    """approximations.py:359-362"""
    mu, L = mvt_unpack(var_param, dim)
    return mu.copy(), df / (df - 2.0) * (L @ L.T)


def mvt_pth_moment(var_param, dim, df, p):
    """approximations.py:364-377"""
    if p not in (2, 4) or not p < df:
        raise ValueError('p = {} is not a supported moment'.format(p))
    _, L = mvt_unpack(var_param, dim)
    w = np.linalg.eigvalsh(L @ L.T)
    c = df / (df - 2.0)
    if p == 2:
        return c * np.sum(w)
    return c ** 2 * (2 * (df - 1.0) / (df - 4.0) * np.sum(w ** 2) + np.sum(w) ** 2)


# --------------------------------------------------------------------------
# Model plugins: log density [S] and its gradient [S,d] at theta[S,d]
# (the reference's Model.__call__, models.py:27-39, evaluates user code; these are
#  the built-in GPU-resident models named by BASELINE.json:north_star.)
# --------------------------------------------------------------------------
def softplus(x):
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def sigmoid(x):
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def gauss_prior(theta, prior_sd):
    d = theta.shape[1]
    lp = -0.5 * np.sum(theta * theta, axis=1) / prior_sd ** 2 - d * math.log(prior_sd * math.sqrt(2 * math.pi))
    return lp, -theta / prior_sd ** 2


def logistic_logp_grad(theta, X, y, prior_sd=10.0, want_grad=True, chunk=65536):
    """f(theta) = -sum_n softplus(-y_n x_n.theta) + log N(theta; 0, prior_sd^2 I), y in {-1,+1}.
    grad = X^T (y * sigmoid(-y * X theta)) - theta/prior_sd^2   (SURVEY.md 8(a) a5, App. A.1)."""
    theta = np.atleast_2d(theta)
    lp, gp = gauss_prior(theta, prior_sd)
    ll = np.zeros(theta.shape[0])
    G = gp.copy() if want_grad else None
    for s in range(0, X.shape[0], chunk):
        Xc, yc = X[s:s + chunk], y[s:s + chunk]
        M = (Xc @ theta.T) * yc[:, None]                # y_n z_ns, [n,S]
        ll -= np.sum(softplus(-M), axis=0)
        if want_grad:
            R = sigmoid(-M) * yc[:, None]
            G += R.T @ Xc
    return lp + ll, G


def probit_logp_grad(theta, X, y, prior_sd=10.0, want_grad=True, chunk=65536):
    """f(theta) = sum_n log Phi(y_n x_n.theta) + Gaussian prior."""
    theta = np.atleast_2d(theta)
    lp, gp = gauss_prior(theta, prior_sd)
    ll = np.zeros(theta.shape[0])
    G = gp.copy() if want_grad else None
    for s in range(0, X.shape[0], chunk):
        Xc, yc = X[s:s + chunk], y[s:s + chunk]
        M = (Xc @ theta.T) * yc[:, None]
        lc = sps.log_ndtr(M)
        ll += np.sum(lc, axis=0)
        if want_grad:
            R = np.exp(-0.5 * M * M - 0.5 * LOG_2PI - lc) * yc[:, None]
            G += R.T @ Xc
    return lp + ll, G


def gauss_target_logp_grad(theta, mean, sd):
    """sum_j norm.logpdf(theta_j; mean_j, sd_j) -- the reference tests' target
    (tests/test_objectives.py:18-19)."""
    theta = np.atleast_2d(theta)
    z = (theta - mean) / sd
    return np.sum(-0.5 * z * z - np.log(sd) - 0.5 * LOG_2PI, axis=1), -z / sd


def student_target_logp_grad(theta, loc, scale, df):
    """sum_j t.logpdf(theta_j; df, loc_j, scale_j) (product Student-t target)."""
    theta = np.atleast_2d(theta)
    z = (theta - loc) / scale
    lp = np.sum(student_t_logpdf(z, df) - np.log(scale), axis=1)
    return lp, -(df + 1.0) * z / ((df + z * z) * scale)


def logistic_hessian(theta, X, y, prior_sd=10.0):
    """Hessian of logistic_logp_grad's f at ONE point: -X^T diag(p(1-p)) X - I/prior_sd^2."""
    a = (X @ theta) * y
    c = sigmoid(a) * sigmoid(-a)
    return -(X * c[:, None]).T @ X - np.eye(theta.size) / prior_sd ** 2


def probit_hessian(theta, X, y, prior_sd=10.0):
    """d^2/da^2 log Phi(a) = -r (a + r), r = phi(a)/Phi(a)."""
    a = (X @ theta) * y
    r = np.exp(-0.5 * a * a - 0.5 * LOG_2PI - sps.log_ndtr(a))
    c = r * (a + r)
    return -(X * c[:, None]).T @ X - np.eye(theta.size) / prior_sd ** 2


def gauss_target_hessian(theta, mean, sd):
    return np.diag(-1.0 / (sd * sd) * np.ones_like(theta))


def student_target_hessian(theta, loc, scale, df):
    z = (theta - loc) / scale
    return np.diag(-(df + 1.0) * (df - z * z) / ((df + z * z) ** 2 * scale * scale))


def hier_linear_layout(G, p):
    """theta = [beta(G*p, group-major), m(p), log_tau, log_sigma]."""
    return G * p + p + 2


def hier_linear_logp_grad(theta, X, y, group, G, p, want_grad=True):
    """Hierarchical linear regression (SURVEY.md 8(d) C4 definition):
    y_i ~ N(x_i . beta_{g(i)}, sigma), beta_g ~ N(m, tau I), m ~ N(0, 10 I),
    log tau ~ N(0,1), log sigma ~ N(0,1); theta unconstrained (log scales, the Jacobian
    is part of the N(0,1) priors on the log scale, i.e. no extra term)."""
    theta = np.atleast_2d(theta)
    S = theta.shape[0]
    beta = theta[:, :G * p].reshape(S, G, p)
    m = theta[:, G * p:G * p + p]
    ltau, lsig = theta[:, -2], theta[:, -1]
    tau, sig = np.exp(ltau), np.exp(lsig)
    N = X.shape[0]
    pred = np.einsum('np,snp->sn', X, beta[:, group, :])
    res = (y[None, :] - pred) / sig[:, None]                       # [S,N]
    lp = np.sum(-0.5 * res * res, axis=1) - N * lsig - 0.5 * N * LOG_2PI
    db = (beta - m[:, None, :]) / tau[:, None, None]               # [S,G,p]
    lp += np.sum(-0.5 * db * db, axis=(1, 2)) - G * p * ltau - 0.5 * G * p * LOG_2PI
    lp += np.sum(-0.5 * (m / 10.0) ** 2, axis=1) - p * math.log(10.0) - 0.5 * p * LOG_2PI
    lp += -0.5 * ltau ** 2 - 0.5 * LOG_2PI - 0.5 * lsig ** 2 - 0.5 * LOG_2PI
    if not want_grad:
        return lp, None
    Gd = np.zeros_like(theta)
    rs = res / sig[:, None]                                        # d/dpred
    gb = np.zeros((S, G, p))
    for g in range(G):
        sel = group == g
        gb[:, g, :] = rs[:, sel] @ X[sel]
    gb -= db / tau[:, None, None]
    Gd[:, :G * p] = gb.reshape(S, G * p)
    Gd[:, G * p:G * p + p] = np.sum(db, axis=1) / tau[:, None] - m / 100.0
    Gd[:, -2] = np.sum(db * db, axis=(1, 2)) - G * p - ltau
    Gd[:, -1] = np.sum(res * res, axis=1) - N - lsig
    return lp, Gd


# --------------------------------------------------------------------------
# Objectives (objectives.py:150-168 ExclusiveKL; :440-463 AlphaDivergence)
# `model(theta) -> (logp[S], grad[S,d])`
# --------------------------------------------------------------------------
def exclusive_kl_meanfield(var_param, base, model, family='gaussian', df=None, path_deriv=False):
    """-ELBO and its gradient for MFGaussian / MFStudentT with injected base draws.
    objectives.py:154-168 + SURVEY.md App. A.1.  Returns (value, grad[2d], logp[S])."""
    S, d = base.shape
    mu, ls = mf_unpack(var_param, d)
    sig = np.exp(ls)
    theta = mu + sig * base
    f, g = model(theta)
    if path_deriv:
        # objectives.py:156-159: mean(f(theta) - log q(theta; stop_grad(lambda)))
        if family == 'gaussian':
            logq = mfg_log_density(var_param, theta)
            g = g + base / sig
        else:
            logq = mft_log_density(var_param, theta, df)
            g = g + (df + 1.0) * base / ((df + base * base) * sig)
        value = -np.mean(f - logq)
        gmu = -np.mean(g, axis=0)
        gls = -np.mean(g * base, axis=0) * sig
    else:
        H = mfg_entropy(var_param, d) if family == 'gaussian' else mft_entropy(var_param, d)
        value = -(np.mean(f) + H)
        gmu = -np.mean(g, axis=0)
        gls = -np.mean(g * base, axis=0) * sig - 1.0
    return value, np.concatenate([gmu, gls]), f


def exclusive_kl_cv_meanfield(var_param, base, model, hessian, method, family='gaussian', df=None, path_deriv=False):
    """ExclusiveKL with the control-variate gradient estimators (objectives.py:170-273, after Miller et al.),
    restated per sample exactly as the reference computes them.  `model(theta[S,d]) -> (f[S], G[S,d])`,
    `hessian(m[d]) -> H[d,d]` (the reference obtains both from autograd).  Returns (value, grad[2d])."""
    S, d = base.shape
    mu, ls = mf_unpack(var_param, d)
    sig = np.exp(ls)
    z = mu + sig * base                                                  # approx.sample (:171)
    if family == 'gaussian':
        m_mean, s_scale = mu, np.sqrt(np.exp(2 * ls))                    # mean_and_cov (:172-173)
    else:
        m_mean, s_scale = mu, np.sqrt(df / (df - 2.0) * np.exp(2 * ls))
    eps = (z - m_mean) / s_scale                                         # :174
    f, dLdm = model(z)
    if path_deriv:                                                       # :176-183
        logq = mfg_log_density(var_param, z) if family == 'gaussian' else mft_log_density(var_param, z, df)
        lower = np.mean(f - logq)
    else:
        H_ent = mfg_entropy(var_param, d) if family == 'gaussian' else mft_entropy(var_param, d)
        lower = np.mean(f) + H_ent
    dLdlns = dLdm * eps * s_scale + 1                                    # :196
    g_hat = np.column_stack([dLdm, dLdlns])
    gm = model(m_mean[None, :])[1][0]                                    # gradient at the mean
    Hm = hessian(m_mean)
    hvps = (s_scale * eps) @ Hm                                          # H symmetric: row s = H (s_scale * eps_s)
    if method == 'full':                                                 # :199-215
        dLdz = gm + hvps
        dLds = dLdz * eps * s_scale + 1.0
        tilde = np.column_stack([dLdz, dLds])
        tilde_mean = np.concatenate([gm, (np.diag(Hm) * s_scale + 1 / s_scale) * s_scale])
        g = np.mean(g_hat - (tilde - tilde_mean), axis=0)
    elif method == 'mean_only':                                          # :216-232
        g_tilde = np.column_stack([gm + hvps, np.zeros_like(hvps)])
        E = np.concatenate([gm, np.zeros(d)])
        g = np.mean(g_hat - (g_tilde - E), axis=0)
    elif method == 'loo_diag_approx':                                    # :233-254
        dLdz = gm + hvps
        dLds = dLdz * (eps * s_scale) + 1
        Hd_sum = np.sum(eps * hvps, axis=0)
        Hd_s = (Hd_sum[None, :] - eps * hvps) / float(S - 1)
        dLds_mu = (Hd_s + 1 / s_scale[None, :]) * s_scale
        g = g_hat.copy()
        g[:, :d] -= hvps
        g[:, d:] -= (dLds - dLds_mu)
        g = np.mean(g, axis=0)
    elif method == 'loo_direct_approx':                                  # :255-268
        dLdz = gm + hvps
        dLds = (dLdz * eps + 1 / s_scale[None, :]) * s_scale
        dLds_mu = (np.sum(dLds, axis=0)[None, :] - dLds) / float(S - 1)
        g = np.mean(g_hat - np.column_stack([hvps, dLds - dLds_mu]), axis=0)
    else:
        raise RuntimeError('Invalid hessian approximation method!')
    return -lower, -g


def alpha_divergence_meanfield(var_param, base, model, alpha, family='gaussian', df=None):
    """objectives.py:443-460; the gradient is NOT divided by mean(scaled)."""
    S, d = base.shape
    mu, ls = mf_unpack(var_param, d)
    sig = np.exp(ls)
    theta = mu + sig * base
    f, g = model(theta)
    logq = mfg_log_density(var_param, theta) if family == 'gaussian' else mft_log_density(var_param, theta, df)
    lw = f - logq
    m = np.max(lw)
    sv = np.exp(lw - m) ** alpha
    value = np.log(np.mean(sv)) / alpha + m
    gmu = alpha / S * (sv @ g)
    gls = alpha / S * (sv @ (g * base * sig + 1.0))
    return value, np.concatenate([gmu, gls]), lw


def _mvt_backprop(Abar, L, w, V):
    """Cotangent of A = sqrtm(L L^T) back to the free parameters (without the entropy /
    log-det term).  SURVEY.md App. A.3."""
    rw = np.sqrt(w)
    M = V.T @ Abar @ V
    Sbar = V @ (M / (rw[:, None] + rw[None, :])) @ V.T
    Lbar = (Sbar + Sbar.T) @ L
    Fbar = np.tril(Lbar, -1) + np.diag(np.diag(Lbar) * np.diag(L))
    return Fbar


def exclusive_kl_mvt(var_param, chi2, z, model, df):
    """objectives.py:154-164 with MultivariateT (entropy branch).  (value, grad, logp)."""
    S, d = z.shape
    mu, L = mvt_unpack(var_param, d)
    A, w, V = sym_sqrt(L @ L.T)
    u = np.sqrt(chi2 / df)
    zu = z / u[:, None]
    theta = mu + zu @ A
    f, g = model(theta)
    value = -(np.mean(f) + mvt_entropy(var_param, d))
    gmu = -np.mean(g, axis=0)
    Abar = -(zu.T @ g) / S
    Fbar = _mvt_backprop(Abar, L, w, V)
    Fbar[np.diag_indices(d)] -= 1.0
    return value, np.concatenate([gmu, Fbar[np.tril_indices(d)]]), f


def alpha_divergence_mvt(var_param, chi2, z, model, df, alpha):
    """objectives.py:443-460 with MultivariateT."""
    S, d = z.shape
    mu, L = mvt_unpack(var_param, d)
    A, w, V = sym_sqrt(L @ L.T)
    u = np.sqrt(chi2 / df)
    zu = z / u[:, None]
    theta = mu + zu @ A
    f, g = model(theta)
    lw = f - mvt_log_density(var_param, theta, df)
    m = np.max(lw)
    sv = np.exp(lw - m) ** alpha
    value = np.log(np.mean(sv)) / alpha + m
    gmu = alpha / S * (sv @ g)
    Abar = alpha / S * (zu.T @ (sv[:, None] * g))
    Fbar = _mvt_backprop(Abar, L, w, V)
    Fbar[np.diag_indices(d)] += alpha / S * np.sum(sv)
    return value, np.concatenate([gmu, Fbar[np.tril_indices(d)]]), lw


# --------------------------------------------------------------------------
# Optimizer steps (optimization.py:188-197 RMSProp, :308-326 Adam)
# --------------------------------------------------------------------------
def rmsprop_direction(state, grad, beta=0.9, jitter=1e-8):
    """state['nu'] starts at grad**2, so step 1 gives nu = grad**2 (optimization.py:189-195)."""
    g2 = grad * grad
    nu = state.get('nu')
    nu = g2.copy() if nu is None else nu
    nu = beta * nu + (1.0 - beta) * g2
    state['nu'] = nu
    return grad / np.sqrt(jitter + nu)


def adam_direction(state, grad, beta1=0.9, beta2=0.999, jitter=1e-8):
    """Reproduces the aliasing quirk (optimization.py:314-320): on the first
    call `momentum` IS the caller's grad array, so `momentum *= beta1` scales grad before the
    `(1-beta1)*grad` term is formed, and grad == momentum when `grad**2` is taken for nu.
    (The caller's array is left untouched here; the arithmetic is what is reproduced.)"""
    if state.get('m') is None:
        g = beta1 * grad                     # grad after the in-place `momentum *= beta1`
        m = g + (1.0 - beta1) * g            # ... and after `momentum += (1-beta1)*grad`
        nu = beta2 * (grad * grad) + (1.0 - beta2) * m * m
        state['m'], state['nu'] = m, nu
        return m / np.sqrt(jitter + nu)
    m = beta1 * state['m'] + (1.0 - beta1) * grad
    nu = beta2 * state['nu'] + (1.0 - beta2) * grad * grad
    state['m'], state['nu'] = m, nu
    return m / np.sqrt(jitter + nu)


# --------------------------------------------------------------------------
# LRGaussian (approximations.py:552-731): var_param = [mu(d), log_sigma(d), B(d,k) row-major]
# (PatternDict insertion order :552-557), Sigma = B B^T + diag(exp(2 log_sigma)).
# Dense d x d algebra on purpose: an independent restatement of what the Woodbury / determinant-lemma
# code of the reference computes.
# --------------------------------------------------------------------------
def lr_unpack(var_param, dim, k):
    vp = np.asarray(var_param, dtype=np.float64)
    return vp[:dim], vp[dim:2 * dim], vp[2 * dim:].reshape(dim, k)


def lr_sigma(var_param, dim, k):
    _, ls, B = lr_unpack(var_param, dim, k)
    return B @ B.T + np.diag(np.exp(2 * ls))


def lr_sample(var_param, z, eps):
    """mu + z B^T + exp(log_sigma) * eps with z drawn FIRST (approximations.py:638-646)."""
    d, k = eps.shape[1], z.shape[1]
    mu, ls, B = lr_unpack(var_param, d, k)
    return mu + z @ B.T + np.exp(ls) * eps


def lr_entropy(var_param, dim, k):
    return 0.5 * dim * (LOG_2PI + 1.0) + 0.5 * np.linalg.slogdet(lr_sigma(var_param, dim, k))[1]


def lr_log_density(var_param, x, k):
    x = np.atleast_2d(x)
    d = x.shape[1]
    mu = lr_unpack(var_param, d, k)[0]
    Sig = lr_sigma(var_param, d, k)
    diff = x - mu
    maha = np.sum(diff * np.linalg.solve(Sig, diff.T).T, axis=1)
    return -0.5 * (d * LOG_2PI + np.linalg.slogdet(Sig)[1] + maha)


def lr_kl(vp0, vp1, dim, k):
    mu0, mu1 = lr_unpack(vp0, dim, k)[0], lr_unpack(vp1, dim, k)[0]
    S0, S1 = lr_sigma(vp0, dim, k), lr_sigma(vp1, dim, k)
    md = mu0 - mu1
    return 0.5 * (np.linalg.slogdet(S1)[1] - np.linalg.slogdet(S0)[1] - dim + md @ np.linalg.solve(S1, md)
                  + np.trace(np.linalg.solve(S1, S0)))


def lr_mean_and_cov(var_param, dim, k):
    return lr_unpack(var_param, dim, k)[0].copy(), lr_sigma(var_param, dim, k)


def lr_pth_moment(var_param, dim, k, p):
    ev = np.linalg.eigvalsh(lr_sigma(var_param, dim, k))
    return np.sum(ev) if p == 2 else 2 * np.sum(ev ** 2) + np.sum(ev) ** 2


def lr_objective(var_param, z, eps, model, kind, alpha=None):
    """ExclusiveKL ('ekl', 'ekl_path'; objectives.py:154-164) and AlphaDivergence ('alpha'; :443-460) for LRGaussian
    with analytic gradients.  Returns (value, grad)."""
    S, d = eps.shape
    k = z.shape[1]
    mu, ls, B = lr_unpack(var_param, d, k)
    sig = np.exp(ls)
    theta = mu + z @ B.T + sig * eps
    f, G = model(theta)
    Sinv = np.linalg.inv(lr_sigma(var_param, d, k))
    u = theta - mu
    a = u @ Sinv                                          # rows: Sigma^-1 u_s
    if kind == 'ekl':
        value = -(np.mean(f) + lr_entropy(var_param, d, k))
        gmu = -np.mean(G, axis=0)
        gls = -(np.mean(G * eps, axis=0) * sig + sig * sig * np.diag(Sinv))
        gB = -(G.T @ z / S + Sinv @ B)
    elif kind == 'ekl_path':
        value = -np.mean(f - lr_log_density(var_param, theta, k))
        gt = G + a
        gmu = -np.mean(gt, axis=0)
        gls = -np.mean(gt * eps, axis=0) * sig
        gB = -gt.T @ z / S
    else:
        lw = f - lr_log_density(var_param, theta, k)
        m = np.max(lw)
        sv = np.exp(lw - m) ** alpha
        value = np.log(np.mean(sv)) / alpha + m
        gmu = alpha / S * (sv @ G)
        dls = G * eps * sig + a * eps * sig + sig * sig * np.diag(Sinv) - sig * sig * a * a
        gls = alpha / S * (sv @ dls)
        gB = np.zeros((d, k))
        for s_ in range(S):
            gB += sv[s_] * (np.outer(G[s_] + a[s_], z[s_]) + Sinv @ B - np.outer(a[s_], a[s_] @ B))
        gB *= alpha / S
    return value, np.concatenate([gmu, gls, gB.reshape(-1)])


# --------------------------------------------------------------------------
# PSIS (_psis.py:113-396) -- selection restatement (no full argsort)

# ---------------------------------------------------------------------------------------------
# NeuralNet / NVPFlow (approximations.py:385-550), forward functions; tanh hidden activations
# ---------------------------------------------------------------------------------------------
def nn_fold(flat, shapes):
    """PatternDict order (:404-407): W0 (C-order), b0, W1, b1, ..."""
    out, off = [], 0
    for a, b in shapes:
        W = flat[off:off + a * b].reshape(a, b)
        off += a * b
        out.append((W, flat[off:off + b]))
        off += b
    assert off == flat.size
    return out


def nn_forward(flat, shapes, x, last='tanh'):
    """NeuralNet.forward (:414-429): tanh layers, `last` in {'tanh', 'identity'}; the log-det term is the
    reference's log|sum_j (f'(out) W^T)_j| with the derivative evaluated at the layer OUTPUT."""
    layers = nn_fold(np.asarray(flat, dtype=np.float64), shapes)
    log_det = np.zeros(x.shape[0])
    for i, (W, b) in enumerate(layers):
        ident = (i + 1 == len(layers)) and last == 'identity'
        x = x @ W + b
        if not ident:
            x = np.tanh(x)
        d = np.ones_like(x) if ident else 1.0 - np.tanh(x) ** 2
        log_det = log_det + np.log(np.abs((d @ W.T).sum(axis=1)))
    return x, log_det


def nvp_split(var_param, shapes_t, shapes_s, n_layers):
    """Flat NVPFlow parameter -> [(t_flat, s_flat)] ("it" before "is", :487-489)."""
    nt = sum(a * b + b for a, b in shapes_t)
    ns = sum(a * b + b for a, b in shapes_s)
    out, off = [], 0
    for _ in range(n_layers):
        out.append((var_param[off:off + nt], var_param[off + nt:off + nt + ns]))
        off += nt + ns
    assert off == var_param.size
    return out


def nvp_g(var_param, shapes_t, shapes_s, mask, z):
    """NVPFlow.g (:493-511)."""
    x = z
    for (tf, sf), m in zip(nvp_split(var_param, shapes_t, shapes_s, len(mask)), mask):
        x_ = x * m
        s = nn_forward(sf, shapes_s, x_, 'tanh')[0] * (1 - m)
        t = nn_forward(tf, shapes_t, x_, 'identity')[0] * (1 - m)
        x = x_ + (1 - m) * (x * np.exp(s) + t)
    return x


def nvp_f(var_param, shapes_t, shapes_s, mask, x):
    """NVPFlow.f (:513-531): (z, log_det_J)."""
    parts = nvp_split(var_param, shapes_t, shapes_s, len(mask))
    log_det, z = np.zeros(x.shape[0]), x
    for i in reversed(range(len(mask))):
        tf, sf = parts[i]
        m = mask[i]
        z_ = m * z
        s = nn_forward(sf, shapes_s, z_, 'tanh')[0] * (1 - m)
        t = nn_forward(tf, shapes_t, z_, 'identity')[0] * (1 - m)
        z = (1 - m) * (z - t) * np.exp(-s) + z_
        log_det = log_det - s.sum(axis=1)
    return z, log_det


def nvp_log_density(var_param, shapes_t, shapes_s, mask, x, prior_param, prior_df=None):
    """NVPFlow.log_density (:533-535) over an MFGaussian (prior_df None) or MFStudentT prior."""
    z, ld = nvp_f(var_param, shapes_t, shapes_s, mask, x)
    lp = mfg_log_density(prior_param, z) if prior_df is None else mft_log_density(prior_param, z, prior_df)
    return lp + ld

# --------------------------------------------------------------------------
def psis_tail_len(n, Reff=1.0):
    """_psis.py:158: M such that cutoff_ind = -M-1."""
    return int(math.ceil(min(0.2 * n, 3.0 * math.sqrt(n / Reff))))


def gpinv(p, k, sigma):
    """Inverse generalised-Pareto CDF for 0<p<1 (_psis.py:335-377, the all-ok branch
    is the only one psislw reaches)."""
    p = np.asarray(p, dtype=np.float64)
    if sigma <= 0:
        return np.full(p.shape, np.nan)
    if abs(k) < np.finfo(float).eps:
        q = -np.log1p(-p)
    else:
        q = np.expm1(-k * np.log1p(-p)) / k
    q = q * sigma
    q = np.where(p == 0, 0.0, q)
    q = np.where(p == 1, np.inf if k >= 0 else -sigma / k, q)
    q = np.where((p < 0) | (p > 1), np.nan, q)
    return q


def gpdfit(x_sorted):
    """Zhang-Stephens empirical-Bayes GPD fit on ASCENDING x (_psis.py:212-332).
    Returns (k, sigma) with the weakly-informative prior applied to k (:323-324)."""
    x = np.asarray(x_sorted, dtype=np.float64)
    n = x.size
    if x.ndim != 1 or n <= 1:
        raise ValueError('Invalid input array.')
    m = 30 + int(math.sqrt(n))
    j = np.arange(1, m + 1, dtype=np.float64) - 0.5
    bs = 1.0 - np.sqrt(m / j)
    bs = bs / (3.0 * x[int(n / 4 + 0.5) - 1])
    bs = bs + 1.0 / x[-1]
    ks = np.mean(np.log1p(-bs[:, None] * x), axis=1)
    with np.errstate(divide='ignore', invalid='ignore'):
        Lj = n * (np.log(-(bs / ks)) - ks - 1.0)
    with np.errstate(over='ignore'):
        w = 1.0 / np.sum(np.exp(Lj - Lj[:, None]), axis=1)
    keep = w >= 10 * np.finfo(float).eps
    w, bs = w[keep], bs[keep]
    w = w / np.sum(w)
    b = np.sum(bs * w)
    k = np.mean(np.log1p(-b * x))
    sigma = -k / b
    k = k * n / (n + 10.0) + 5.0 / (n + 10.0)
    return k, sigma


def sumlogs(x):
    """log(sum(exp(x))) (_psis.py:380-396)."""
    mx = np.max(x)
    return np.log(np.sum(np.exp(x - mx))) + mx


def psislw_1d(lw, Reff=1.0, return_tail=False):
    """One column of _psis.py:163-203.  O(n) selection for the cutoff instead of the full
    argsort (:168): the cutoff is an order-statistic VALUE and the tail is {x > cutoff},
    so the result is identical.  Tail ties are ranked by index
    (the reference's argsort(x2) uses an unstable sort)."""
    x = np.array(lw, dtype=np.float64, copy=True)
    n = x.size
    if n <= 1:
        raise ValueError('More than one log-weight needed.')
    M = psis_tail_len(n, Reff)
    x -= np.max(x)
    kth = n - M - 1                                   # index of the (M+1)-th largest
    xcut = max(np.partition(x, kth)[kth], math.log(np.finfo(float).tiny))
    expcut = math.exp(xcut)
    tail = np.flatnonzero(x > xcut)
    n2 = tail.size
    order = None
    if n2 <= 4:
        k = np.inf
    else:
        x2 = x[tail]
        order = np.argsort(x2, kind='stable')
        k, sigma = gpdfit(np.exp(x2[order]) - expcut)
    if k >= 1.0 / 3.0 and not np.isinf(k):
        q = gpinv((np.arange(n2) + 0.5) / n2, k, sigma) + expcut
        x[tail[order]] = np.log(q)
        x[x > 0] = 0.0
    x -= sumlogs(x)
    if return_tail:
        return x, k, tail, (tail[order] if order is not None else tail)
    return x, k


def psislw_argsort_1d(lw, Reff=1.0):
    """One column of _psis.py:163-203 with the reference's own COST STRUCTURE: a full argsort of the n
    log-weights (:168), the cutoff read through the sort permutation (:170-173), a second argsort of the
    tail (:184).  Same results as psislw_1d (checked in tests/test_oracle_golden.py); this is the variant
    bench.py times as the CPU baseline of the PSIS leg, because it does the work the reference does."""
    x = np.array(lw, dtype=np.float64, copy=True)
    n = x.size
    if n <= 1:
        raise ValueError('More than one log-weight needed.')
    cut_pos = -psis_tail_len(n, Reff) - 1
    x -= np.max(x)
    perm = np.argsort(x)
    xcut = max(x[perm[cut_pos]], math.log(np.finfo(float).tiny))
    expcut = math.exp(xcut)
    tail = np.flatnonzero(x > xcut)
    n2 = tail.size
    k, order = np.inf, None
    if n2 > 4:
        x2 = x[tail]
        order = np.argsort(x2)
        k, sigma = gpdfit(np.exp(x2[order]) - expcut)
    if k >= 1.0 / 3.0 and not np.isinf(k):
        q = gpinv((np.arange(n2) + 0.5) / n2, k, sigma) + expcut
        x[tail[order]] = np.log(q)
        x[x > 0] = 0.0
    x -= sumlogs(x)
    return x, k


def psis_shard_record(x_local, idx_off, M):
    """Draw-sharded restatement of _psis.py:163-203, stage 1 (csrc/psis.cu
    psis_export_kernel): what one rank ships.  [local max, c_r, log-sum-exp (relative to the local
    max) of the draws not represented, count], the rank's top M+1 values -- values above its
    (M+1)-th largest c_r, padded with copies of c_r -- and their global indices (-1 = padding)."""
    x = np.asarray(x_local, dtype=np.float64)
    K = M + 1
    vals, idx = np.full(K, -np.inf), np.full(K, -1, dtype=np.int64)
    mx = x.max()
    if x.size < K:
        vals[:x.size], idx[:x.size] = x, idx_off + np.arange(x.size)
        return np.array([mx, -np.inf, -np.inf, x.size]), vals, idx
    c = np.partition(x, x.size - K)[x.size - K]
    tail = np.flatnonzero((x > c) & (x - mx > math.log(np.finfo(float).tiny)))
    vals[:tail.size], idx[:tail.size] = x[tail], idx_off + tail
    vals[tail.size:] = c
    rest = np.ones(x.size, dtype=bool)
    rest[tail] = False
    body = np.exp(x[rest] - mx).sum() - (K - tail.size) * math.exp(c - mx)
    return np.array([mx, c, math.log(body) if body > 0 else -np.inf, K]), vals, idx


def psis_merge_records(records, M, n_global):
    """Stage 2 of the sharded restatement: from every rank's record, the global maximum, the
    shifted cutoff (_psis.py:170-173), the tail (global indices, shifted values) and the
    log-sum-exp of everything that is not in the tail (relative to the global maximum)."""
    heads = np.stack([r[0] for r in records])
    vals = np.concatenate([r[1][:int(r[0][3])] for r in records])
    idx = np.concatenate([r[2][:int(r[0][3])] for r in records])
    mx = heads[:, 0].max()
    c = np.partition(vals, vals.size - (M + 1))[vals.size - (M + 1)]          # (M+1)-th largest of the union
    cutoff = max(c - mx, math.log(np.finfo(float).tiny))
    v = vals - mx
    in_tail = v > cutoff
    body = np.exp(v[~in_tail]).sum()
    for h in heads:
        if h[2] > -np.inf:
            body += math.exp(h[2] + h[0] - mx)
    return mx, cutoff, idx[in_tail], v[in_tail], body


def psislw_sharded(parts, Reff=1.0):
    """_psis.py:163-203 on the concatenation of `parts`, computed the draw-sharded way; returns
    the list of smoothed, normalised parts and k-hat.  Must equal psislw_1d(np.concatenate(parts))."""
    sizes = [len(p) for p in parts]
    n = sum(sizes)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    M = psis_tail_len(n, Reff)
    recs = [psis_shard_record(p, offs[r], M) for r, p in enumerate(parts)]
    mx, cutoff, tidx, tv, body = psis_merge_records(recs, M, n)
    expcut = math.exp(cutoff)
    order = np.lexsort((tidx, tv))
    tidx, tv = tidx[order], tv[order]
    n2 = tv.size
    k = np.inf
    smoothed = None
    if n2 > 4:
        k, sigma = gpdfit(np.exp(tv) - expcut)
        if k >= 1.0 / 3.0 and not np.isinf(k):
            smoothed = np.minimum(np.log(gpinv((np.arange(n2) + 0.5) / n2, k, sigma) + expcut), 0.0)
    tail_out = smoothed if smoothed is not None else tv
    lse = math.log(body + np.exp(tail_out).sum())
    outs = []
    for r, p in enumerate(parts):
        o = np.asarray(p, dtype=np.float64) - mx
        mine = (tidx >= offs[r]) & (tidx < offs[r + 1])
        o[tidx[mine] - offs[r]] = tail_out[mine]
        outs.append(o - lse)
    return outs, k


def psislw(lw, Reff=1.0):
    """_psis.py:113-209 for 1-D or [n,m] input (each column separately)."""
    lw = np.asarray(lw, dtype=np.float64)
    if lw.ndim == 1:
        return psislw_1d(lw, Reff)
    if lw.ndim != 2:
        raise ValueError('Argument `lw` must be 1 or 2 dimensional.')
    out = np.empty(lw.shape, order='F')
    ks = np.empty(lw.shape[1])
    for i in range(lw.shape[1]):
        out[:, i], ks[i] = psislw_1d(lw[:, i], Reff)
    return out, ks


def psisloo(log_lik, Reff=1.0):
    """_psis.py:69-110: PSIS leave-one-out from n x m log-likelihood draws.  Returns (loo, loos[m], ks[m])."""
    ll = np.asarray(log_lik, dtype=np.float64)
    lw, ks = psislw(-ll, Reff)
    lw = lw + ll
    loos = np.array([sumlogs(lw[:, i]) for i in range(lw.shape[1])])
    return loos.sum(), loos, ks


# --------------------------------------------------------------------------
# Divergence / Wasserstein / error bounds (diagnostics.py:13-219)
# --------------------------------------------------------------------------
def divergence_bound(lw, alpha=2.0, log_norm_bound=None):
    """diagnostics.py:148-186.  Returns (d_alpha, log_norm_bound, cubo)."""
    if alpha <= 1:
        raise ValueError('alpha must be greater than 1')
    lw = np.asarray(lw, dtype=np.float64)
    mx = np.max(lw)
    cubo = np.log(np.mean(np.exp(lw - mx) ** alpha)) / alpha + mx
    if log_norm_bound is None:
        log_norm_bound = np.mean(lw)
    return alpha / (alpha - 1.0) * (cubo - log_norm_bound), log_norm_bound, cubo


def wasserstein_bounds(d2, samples=None, moment_bound_fn=None):
    """diagnostics.py:106-145; sample moments are per-coordinate central power sums."""
    if moment_bound_fn is None:
        if samples is None:
            raise ValueError('must provides samples if moment_bound_fn not given')
        x = np.asarray(samples, dtype=np.float64)
        if x.ndim == 1:
            x = x[:, None]
        c = x - np.mean(x, axis=0, keepdims=True)

        def moment_bound_fn(p):
            return np.mean(np.sum(c ** p, axis=1))
    out = {}
    for p in (1, 2):
        Cp = moment_bound_fn(2 * p)
        out['W%d' % p] = 2 * Cp ** (0.5 / p) * np.expm1(d2) ** (0.5 / p)
    return out


def error_bounds(W1=np.inf, W2=np.inf, q_var=np.inf, p_var=np.inf):
    """diagnostics.py:73-103, :213-219"""
    def nrm(v):
        return np.linalg.norm(v, ord=2) if np.asarray(v).ndim == 2 else v
    qv, pv = nrm(q_var), nrm(p_var)
    min_var = qv if pv is None else np.min([qv, pv], axis=0)
    return dict(mean_error=min(W1, W2), std_error=W2,
                cov_error=2 * (np.sqrt(min_var) * W2 + W2 ** 2))


def all_diagnostics(lw, samples=None, moment_bound_fn=None, q_var=None, p_var=None,
                    log_norm_bound=None):
    """diagnostics.py:13-64"""
    d2, lnb, _ = divergence_bound(lw, log_norm_bound=log_norm_bound)
    res = wasserstein_bounds(d2, samples=samples, moment_bound_fn=moment_bound_fn)
    if q_var is None and samples is not None:
        q_var = np.cov(np.asarray(samples).T)
    res.update(error_bounds(q_var=q_var, p_var=p_var, **res))
    res['d2'] = d2
    res['log_norm_bound'] = lnb
    return res


# --------------------------------------------------------------------------
# One full ELBO-gradient iteration (the bench "step") for the CPU baseline
# --------------------------------------------------------------------------
def elbo_step_logistic(var_param, eps, X, y, opt_state, lr=0.01, prior_sd=10.0):
    """objective(var_param) -> descent_direction -> update, as the reference loop does
    (optimization.py:95-98)."""
    value, grad, _ = exclusive_kl_meanfield(
        var_param, eps, lambda th: logistic_logp_grad(th, X, y, prior_sd))
    direction = rmsprop_direction(opt_state, grad)
    return var_param - lr * direction, value, grad
