"""TEST INFRASTRUCTURE ONLY.  `autograd.scipy.linalg.sqrtm` stand-in.

Forward: scipy.linalg.sqrtm (what autograd wraps).  Backward: the Sylvester
equation A*Xb + Xb*A = G that autograd's own sqrtm VJP solves, done here in the
eigenbasis of the symmetric root (valid for the PSD inputs viabel passes)."""
import numpy as _np
import scipy.linalg as _sl
import torch as _torch

from .._box import Box, is_box


class _Sqrtm(_torch.autograd.Function):
    @staticmethod
    def forward(ctx, S):
        A = _np.real(_sl.sqrtm(S.detach().numpy()))
        A = _torch.from_numpy(_np.ascontiguousarray(A))
        ctx.save_for_backward(A)
        return A

    @staticmethod
    def backward(ctx, G):
        (A,) = ctx.saved_tensors
        w, V = _torch.linalg.eigh(0.5 * (A + A.T))
        M = V.T @ G @ V
        X = M / (w[:, None] + w[None, :])
        return V @ X @ V.T


def sqrtm(S):
    if is_box((S,)):
        return Box(_Sqrtm.apply(S.t))
    return _sl.sqrtm(S)
