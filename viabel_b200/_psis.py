"""Pareto smoothed importance sampling -- drop-in mirror of viabel/_psis.py (psislw :113-209).

The n log-weights stay on the device; the reference's full argsort is replaced by a
threshold + radix-select pipeline (see csrc/psis.cu).  gpdfitnew / gpinv / sumlogs keep their
reference signatures as thin host helpers for small arrays.
"""
import numpy as np
import torch

from . import _lib
from ._tensor import F64, device, is_host, to_dev

__all__ = ['psislw', 'psislw_device', 'gpdfitnew', 'gpinv', 'sumlogs']

R_KHAT, R_SIGMA, R_N2, R_CUTOFF, R_LSE, R_MAX, R_STATUS, R_SUMV, R_SUMEXP2V, R_M, R_NCAND, R_SMOOTHED = range(12)

_ws_cache = {}


def _workspace(n, reff):
    key = (int(n), float(reff), torch.cuda.current_device())
    ws = _ws_cache.get(key)
    if ws is None:
        nbytes = _lib.lib.vb_psis_workspace_bytes(n, reff)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device())
        if len(_ws_cache) > 8:
            _ws_cache.clear()
        _ws_cache[key] = ws
    return ws


def psislw_device(lw, out=None, Reff=1.0, want_tail=False, exact=False):
    """PSIS of one column of log-weights held in a contiguous CUDA float64 tensor.

    Enqueues the whole pipeline without a host sync.  Returns (out, result, tail_idx, tail_rank)
    where result is the 16-slot device vector described in include/viabel_b200.h.  `out` may be
    `lw` itself (in place) or None (k-hat only)."""
    n = lw.numel()
    if n <= 1:
        raise ValueError("More than one log-weight needed.")
    result = torch.empty(16, dtype=F64, device=lw.device)
    tail_idx = tail_rank = None
    if want_tail:
        cap = _lib.lib.vb_psis_tail_capacity(n, Reff)
        tail_idx = torch.empty(cap, dtype=torch.int64, device=lw.device)
        tail_rank = torch.empty(cap, dtype=torch.int32, device=lw.device)
    ws = _workspace(n, Reff)
    _lib.check(_lib.lib.vb_psislw_f64(_lib.ptr(lw), _lib.ptr(out), n, float(Reff), int(exact),
                                      _lib.ptr(result), _lib.ptr(tail_idx), _lib.ptr(tail_rank),
                                      _lib.ptr(ws), ws.numel(), _lib.stream()))
    return out, result, tail_idx, tail_rank


def _psis_column(col, out, Reff, want_tail):
    o, result, ti, tr = psislw_device(col, out, Reff, want_tail)
    res = result.cpu().numpy()
    if res[R_STATUS] == 1:          # sampled threshold missed the tail: exact radix passes
        o, result, ti, tr = psislw_device(col, out, Reff, want_tail, exact=True)
        res = result.cpu().numpy()
    if res[R_STATUS] != 0:
        raise RuntimeError('viabel_b200: PSIS failed with status %d' % int(res[R_STATUS]))
    return res, ti, tr


def psislw(lw, Reff=1.0, overwrite_lw=False, return_tail=False):
    """Pareto smoothed importance sampling (_psis.py:113-209).

    lw: array of n log weights, or n x m for m sets (numpy array or CUDA tensor).
    Returns (lw_out, kss) like the reference: smoothed, normalised log weights and the Pareto
    tail index/indices.  With overwrite_lw=True a CUDA tensor (or F-contiguous numpy array) is
    smoothed in place.  return_tail=True (1-D only) appends (tail_idx, tail_rank)."""
    host = is_host(lw)
    if host:
        lw_np = np.asarray(lw)
        ndim = lw_np.ndim
    else:
        ndim = lw.dim()
    if ndim not in (1, 2):
        raise ValueError("Argument `lw` must be 1 or 2 dimensional.")
    n = lw_np.shape[0] if host else lw.shape[0]
    if n <= 1:
        raise ValueError("More than one log-weight needed.")

    if ndim == 1:
        col = to_dev(lw_np) if host else lw.to(F64).contiguous()
        inplace = (not host) and overwrite_lw and col.data_ptr() == lw.data_ptr()
        out = col if (inplace or host) else torch.empty_like(col)
        res, ti, tr = _psis_column(col, out, Reff, return_tail)
        k = float(res[R_KHAT])
        if host:
            out_np = out.cpu().numpy()
            if overwrite_lw and isinstance(lw, np.ndarray) and lw.flags.f_contiguous and lw.dtype == np.float64:
                lw[...] = out_np
                out_np = lw
            ret = (out_np, k)
        else:
            ret = (out, k)
        if return_tail:
            n2 = int(res[R_N2])
            ret = ret + (ti[:n2].cpu().numpy() if host else ti[:n2], tr[:n2].cpu().numpy() if host else tr[:n2])
        return ret

    if return_tail:
        raise ValueError('return_tail is only available for 1-D input')
    m = lw_np.shape[1] if host else lw.shape[1]
    kss = np.empty(m)
    if host:
        out_np = lw_np if (overwrite_lw and lw_np.flags.f_contiguous and lw_np.dtype == np.float64) \
            else np.copy(lw_np, order='F').astype(np.float64, copy=False)
        for i in range(m):
            col = to_dev(np.ascontiguousarray(lw_np[:, i]))
            res, _, _ = _psis_column(col, col, Reff, False)
            out_np[:, i] = col.cpu().numpy()
            kss[i] = res[R_KHAT]
        return out_np, kss
    cols = lw.to(F64).t().contiguous()           # [m, n]: each set contiguous
    for i in range(m):
        res, _, _ = _psis_column(cols[i], cols[i], Reff, False)
        kss[i] = res[R_KHAT]
    out = cols.t()
    if overwrite_lw:
        lw.copy_(out)
        out = lw
    return out, kss


def gpinv(p, k, sigma):
    """Inverse generalised Pareto distribution function (_psis.py:335-377)."""
    p = np.asarray(p, dtype=np.float64)
    x = np.full(p.shape, np.nan)
    if sigma <= 0:
        return x
    ok = (p > 0) & (p < 1)
    if abs(k) < np.finfo(float).eps:
        x[ok] = -np.log1p(-p[ok])
    else:
        x[ok] = np.expm1(-k * np.log1p(-p[ok])) / k
    x *= sigma
    x[p == 0] = 0
    x[p == 1] = np.inf if k >= 0 else -sigma / k
    return x


def gpdfitnew(x, sort=True, sort_in_place=False, return_quadrature=False):
    """Zhang-Stephens estimate of the generalised Pareto parameters (_psis.py:212-332); host
    helper for small arrays -- psislw runs the same fit on the device."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 1 or len(x) <= 1:
        raise ValueError("Invalid input array.")
    if sort is True:
        if sort_in_place:
            x.sort()
            xs = x
        else:
            xs = np.sort(x)
    elif sort is False:
        xs = x
    else:
        xs = x[sort]
    n = len(xs)
    m = 30 + int(np.sqrt(n))
    j = np.arange(1, m + 1, dtype=float) - 0.5
    bs = (1 - np.sqrt(m / j)) / (3 * xs[int(n / 4 + 0.5) - 1]) + 1 / xs[-1]
    ks = np.mean(np.log1p(-bs[:, None] * xs), axis=1)
    with np.errstate(all='ignore'):
        L = n * (np.log(-bs / ks) - ks - 1)
        w = 1 / np.sum(np.exp(L - L[:, None]), axis=1)
    keep = w >= 10 * np.finfo(float).eps
    w, bsk = w[keep], bs[keep]
    w = w / w.sum()
    b = np.sum(bsk * w)
    k = np.mean(np.log1p(-b * xs))
    sigma = -k / b
    k = k * n / (n + 10) + 5 / (n + 10)
    if return_quadrature:
        kq = np.mean(np.log1p(-bsk[:, None] * xs), axis=1) * n / (n + 10) + 5 / (n + 10)
        return k, sigma, kq, w
    return k, sigma


def sumlogs(x, axis=None, out=None):
    """log(sum(exp(x))) (_psis.py:380-396)."""
    x = np.asarray(x)
    mx = x.max(axis=axis, keepdims=True)
    r = np.log(np.sum(np.exp(x - mx), axis=axis)) + np.squeeze(mx)
    if out is not None:
        out[...] = r
        return out
    return r
