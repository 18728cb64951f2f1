"""CPU emulation (numpy + torch float8 dtypes, no GPU) of an operand scheme for the fast sweep's GEMM1 in which the two
correction passes run on fp8 operands:  z = Xh.Th [fp16 x fp16]  +  Xl8.Th8  +  Xh8.Tl8  [scaled fp8 x fp8],
against the current three fp16 passes and against dropping a correction pass.  Prints the relative errors of
(ll, gmu, ge) at a near-converged point and at init for a logistic problem of the bench's kind."""
import numpy as np
import torch


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def fp8(a, kind, scale):
    """round a*scale to float8 (saturating cast of torch), return the value / scale as float32"""
    dt = torch.float8_e4m3fn if kind == 'e4m3' else torch.float8_e5m2
    t = torch.from_numpy(np.ascontiguousarray(a * scale, dtype=np.float32))
    return (t.to(dt).to(torch.float32).numpy() / np.float32(scale)).astype(np.float32)


def split16(a):
    hi = a.astype(np.float16)
    lo = (a - hi.astype(np.float64)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def main(N=65536, d=512, S=64, seed=3):
    rs = np.random.RandomState(seed)
    X = rs.randn(N, d)
    beta = rs.randn(d) / np.sqrt(d)
    y = np.where(rs.rand(N) < 1 / (1 + np.exp(-X @ beta)), 1.0, -1.0)
    base = rs.randn(S, d).astype(np.float16).astype(np.float64)
    Xy = X * y[:, None]
    for name, theta in (('near-converged', beta + np.exp(-3.5) * base), ('init', 0.0 + np.exp(2.0) * base)):
        a = Xy @ theta.T
        ll0 = -np.logaddexp(0, -a).sum(axis=0)
        R0 = 1 / (1 + np.exp(a))
        gmu0 = Xy.T @ R0.sum(axis=1)
        ge0 = (Xy * (R0 @ base)).sum(axis=0)

        def sweep(z32, Xe2=None):
            Xg = Xy if Xe2 is None else Xe2
            z = z32.astype(np.float32)
            t = np.exp2(-np.abs(z) * np.float32(1.4426950408889634))
            sp = np.maximum(-z, 0) + np.log2(1 + t) * np.float32(0.6931471805599453)
            r = (np.where(z >= 0, t, np.float32(1.0)) / (1 + t)).astype(np.float32)
            ll = -sp.astype(np.float64).sum(axis=0)
            r16 = r.astype(np.float16).astype(np.float32)
            T = r16 @ base.astype(np.float32)
            ge = (Xg * T.astype(np.float64)).sum(axis=0)
            gmu = Xg.T @ r.astype(np.float64).sum(axis=1)
            return max(relerr(ll, ll0), relerr(gmu, gmu0), relerr(ge, ge0)), relerr(gmu, gmu0), relerr(ge, ge0)

        Xh, Xl = split16(Xy)
        Th, Tl = split16(theta)
        z3 = Xh @ Th.T + Xl @ Th.T + Xh @ Tl.T
        z2 = Xh @ Th.T + Xh @ Tl.T                      # X_l.Theta_h dropped
        print('%s:  three fp16 passes %.2e   without Xl.Th %.2e' % (name, sweep(z3)[0], sweep(z2)[0]))
        # RECIPROCAL static scales (one fp32 accumulator: the product of a pass's two operand scales must be 1), all four
        # operands e5m2 (30 binades: range-safe): X_l * 2^ax with Theta_h * 2^-ax, and X_h * 2^-b with Theta_l * 2^b
        for ax, b in ((2, 8), (0, 6), (4, 10), (2, 4)):
            for kx in ('e5m2', 'e4m3'):
                Xl8 = fp8(Xl, kx, 2.0 ** ax)
                Th8 = fp8(Th, 'e5m2', 2.0 ** -ax)
                Xh8 = fp8(Xh, 'e5m2', 2.0 ** -b)
                Tl8 = fp8(Tl, 'e5m2', 2.0 ** b)
                z = Xh @ Th.T + Xl8 @ Th8.T + Xh8 @ Tl8.T
                e = sweep(z)
                print('   fp8 corrections, reciprocal scales ax=%d b=%d, X_l as %s: max %.2e  gmu %.2e  ge %.2e' % (ax, b, kx, e[0], e[1], e[2]))
                e = sweep(z, Xh.astype(np.float64) + Xl8.astype(np.float64))
                print('      + E2 on X_h + X_l8:                                        max %.2e  gmu %.2e  ge %.2e' % e)
                e = sweep(z3, Xh.astype(np.float64))
                print('      (three fp16 passes, E2 on X_h alone:                        max %.2e  gmu %.2e  ge %.2e)' % e)
        # E2 with X = X_h + X_l8 instead of X_h + X_l (the gradient's X factor)


if __name__ == '__main__':
    main()
