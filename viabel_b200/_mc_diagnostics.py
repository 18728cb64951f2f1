"""Convergence statistics for FASO / RAABBVI -- mirror of viabel/_mc_diagnostics.py
(autocov :7-37, ess :40-99, MCSE :102-121, compute_R_hat :124-160, R_hat_convergence_check
:163-184).  Host numpy, but batched over all parameters (the reference loops over them in
Python, :119)."""
import numpy as np
from scipy.fft import next_fast_len

__all__ = ['autocov', 'ess', 'MCSE', 'compute_R_hat', 'R_hat_convergence_check']


def autocov(samples, axis=-1):
    """Autocovariance at every lag via FFT (same normalisation as the reference: / n)."""
    x = np.asarray(samples, dtype=np.float64)
    axis = axis if axis >= 0 else x.ndim + axis
    n = x.shape[axis]
    m = next_fast_len(2 * n)
    x = x - x.mean(axis=axis, keepdims=True)
    f = np.fft.rfft(x, n=m, axis=axis)
    cov = np.fft.irfft(f * np.conjugate(f), n=m, axis=axis)
    sl = [slice(None)] * x.ndim
    sl[axis] = slice(0, n)
    return cov[tuple(sl)] / n


def _ess_from_acov(acov, n_chain=1):
    """Geyer initial positive / monotone sequence estimator on one chain's autocovariance
    (reference ess(), :56-99)."""
    n_draw = acov.shape[0]
    mean_var = acov[0] * n_draw / (n_draw - 1.0)
    var_plus = mean_var * (n_draw - 1.0) / n_draw
    rho = np.zeros(n_draw)
    even = 1.0
    rho[0] = even
    odd = 1.0 - (mean_var - acov[1]) / var_plus
    rho[1] = odd
    t = 1
    while t < (n_draw - 3) and (even + odd) > 0.0:
        even = 1.0 - (mean_var - acov[t + 1]) / var_plus
        odd = 1.0 - (mean_var - acov[t + 2]) / var_plus
        if (even + odd) >= 0:
            rho[t + 1] = even
            rho[t + 2] = odd
        t += 2
    max_t = t - 2
    if even > 0:
        rho[max_t + 1] = even
    t = 1
    while t <= max_t - 2:
        if (rho[t + 1] + rho[t + 2]) > (rho[t - 1] + rho[t]):
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    total = n_chain * n_draw          # the reference keeps going for n_chain > 1 (its ValueError is never raised)
    tau = -1.0 + 2.0 * np.sum(rho[:max_t + 1]) + np.sum(rho[max_t + 1:max_t + 2])
    tau = max(tau, 1 / np.log10(total))
    out = total / tau
    if np.isnan(rho).any():
        out = np.nan
    return out


def ess(samples):
    """Effective sample size of a (n_chains, n_iters) array (FASO passes one chain)."""
    samples = np.asarray(samples, dtype=np.float64)
    acov = autocov(samples, axis=1)
    return _ess_from_acov(np.mean(acov, axis=0), samples.shape[0])


def MCSE(sample):
    """Monte Carlo standard error per column of sample[n_iters, P]; one batched FFT."""
    sample = np.asarray(sample, dtype=np.float64)
    n_iters, d = sample.shape
    sd_dev = np.sqrt(np.var(sample, ddof=1, axis=0))
    with np.errstate(all='ignore'):
        acov = autocov(sample.T, axis=1)                 # [P, n_iters]
        eff = [_ess_from_acov(acov[i]) for i in range(d)]
        mcse = sd_dev / np.sqrt(eff)
    return eff, mcse


def compute_R_hat(chains, warmup=0, jitter=1e-8):
    """Split-R-hat of one chain (:124-160)."""
    chains = np.asarray(chains, dtype=np.float64)[warmup:, :]
    n_iters, d = chains.shape
    if n_iters % 2 == 1:
        n_iters -= 1
        chains = chains[:n_iters, :]
    half = n_iters // 2
    psi = chains.reshape(2, half, d)
    means = psi.mean(axis=1)
    grand = means.mean(axis=0)
    s2 = np.sum((psi - means[:, None, :]) ** 2, axis=1) / (half - 1)
    B = half * np.sum((means - grand) ** 2, axis=0) / (2 - 1)
    W = np.nanmean(s2, axis=0) + jitter
    return np.sqrt((half - 1) / half + B / (half * W))


def R_hat_convergence_check(samples, windows, Rhat_threshold=1.1):
    """(:163-184) samples: array-like [n_iters, P]; returns (success, best window)."""
    samples = np.asarray(samples)
    vals = [np.max(compute_R_hat(samples[-int(w):], 0)) for w in windows]
    best = int(np.argmin(vals))
    return vals[best] <= Rhat_threshold, windows[best]
