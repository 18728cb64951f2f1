"""Convergence statistics for FASO / RAABBVI -- mirror of viabel/_mc_diagnostics.py
(autocov :7-37, ess :40-99, MCSE :102-121, compute_R_hat :124-160, R_hat_convergence_check
:163-184).  Host numpy, but batched over all parameters (the reference loops over them in
Python, :119)."""
import numpy as np
from scipy.fft import next_fast_len

__all__ = ['autocov', 'ess', 'MCSE', 'compute_R_hat', 'R_hat_convergence_check', 'RingStats']


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


class RingStats(object):
    """The same statistics computed ON THE DEVICE over the iterate ring the fused step writes (engine.FusedStep.
    param_hist), batched over all parameters: nothing but a handful of scalars (R-hat per window) or two
    P-vectors (ESS, MCSE) ever crosses to the host.  Kernels: csrc/faso.cu; the FFT is cuFFT through torch.fft."""

    def __init__(self, hist, ring, P):
        import torch
        self.torch = torch
        self.hist, self.ring, self.P = hist, int(ring), int(P)

    def rhat_max(self, end, windows, jitter=1e-8):
        """max over parameters of the split-R-hat of the last W iterates, one value per window (host array)."""
        import ctypes
        from . import _lib
        torch = self.torch
        wins = [int(w) for w in windows]
        out = torch.empty(len(wins), dtype=torch.float64, device=self.hist.device)
        ws = torch.empty(max(8, _lib.lib.vb_faso_rhat_workspace_bytes(self.P, len(wins))), dtype=torch.uint8,
                         device=self.hist.device)
        arr = (ctypes.c_int64 * len(wins))(*wins)
        _lib.check(_lib.lib.vb_faso_rhat_f64(_lib.ptr(self.hist), self.ring, self.P, int(end), arr, len(wins), float(jitter),
                                             _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream()))
        return out.cpu().numpy()

    def convergence_check(self, end, windows, Rhat_threshold=1.1):
        """R_hat_convergence_check (:163-184) on the ring."""
        vals = self.rhat_max(end, windows)
        best = int(np.argmin(vals))
        return bool(vals[best] <= Rhat_threshold), windows[best]

    def window_mean(self, end, W, want_css=False):
        """Column means of the last W iterates (device tensor); with want_css also sum (x - mean)^2."""
        from . import _lib
        torch = self.torch
        mean = torch.empty(self.P, dtype=torch.float64, device=self.hist.device)
        css = torch.empty(self.P, dtype=torch.float64, device=self.hist.device) if want_css else None
        _lib.check(_lib.lib.vb_ring_mean_f64(_lib.ptr(self.hist), self.ring, self.P, int(end), int(W), _lib.ptr(mean),
                                             _lib.ptr(css), _lib.stream()))
        return (mean, css) if want_css else mean

    def mcse(self, end, W):
        """(ess[P], mcse[P], mean[P]) of the last W iterates as host arrays (MCSE, :102-121)."""
        from . import _lib
        torch = self.torch
        W = int(W)
        mean, css = self.window_mean(end, W, want_css=True)
        m = int(next_fast_len(2 * W))
        centered = torch.empty(m, self.P, dtype=torch.float64, device=self.hist.device)
        _lib.check(_lib.lib.vb_faso_center_f64(_lib.ptr(self.hist), self.ring, self.P, int(end), W, m, _lib.ptr(mean),
                                               _lib.ptr(centered), _lib.stream()))
        f = torch.fft.rfft(centered, dim=0)
        acov = torch.fft.irfft(f * f.conj(), n=m, dim=0).contiguous()          # [m, P], unnormalised
        ess_d = torch.empty(self.P, dtype=torch.float64, device=self.hist.device)
        _lib.check(_lib.lib.vb_faso_ess_f64(_lib.ptr(acov), self.P, 1.0 / W, W, self.P, _lib.ptr(ess_d), _lib.stream()))
        sd = torch.sqrt(css / (W - 1.0))
        out = torch.stack([ess_d, sd / torch.sqrt(ess_d), mean]).cpu().numpy()
        return out[0], out[1], out[2]
