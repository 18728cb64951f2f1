import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import viabel_b200 as vb
lib, ptr, st = vb._lib.lib, vb._lib.ptr, vb._lib.stream
G, p, n_per, S = 65, 31, 200, 256
rs = np.random.RandomState(20260118)
N = G * n_per
group = np.repeat(np.arange(G), n_per)
X = rs.randn(N, p); m = rs.randn(p); beta = m + 0.5 * rs.randn(G, p)
y = np.sum(X * beta[group], axis=1) + 0.3 * rs.randn(N)
model = vb.HierarchicalLinearRegression(X, y, group, G)
d = model.dim
approx = vb.MultivariateT(d, 100, seed=1234)
vp0 = approx.init_param()
F = np.zeros((d, d)); F[np.diag_indices(d)] = 0.5 * np.log(0.01)
vp0[d:] = F[np.tril_indices(d)]
vp0[:G * p] = beta.reshape(-1); vp0[G * p:G * p + p] = m
vp = torch.as_tensor(vp0, device='cuda')
L, hl, w, V = approx.decompose(vp)
chi2, z = approx.base_draws(S)
theta, P, zu2 = approx.transform(vp, chi2, z, w, V)
f, Gm = model.logp_and_grad(theta)
fin = lambda t: bool(torch.isfinite(t).all())
print('inputs finite', fin(L), fin(hl), fin(w), fin(V), fin(P), fin(zu2), fin(f), fin(Gm), 'hl', float(hl))
out = torch.full((1 + vp.numel(),), 7.0, dtype=torch.float64, device='cuda')
nbytes = lib.vb_mvt_objective_workspace_bytes(S, d)
ws = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
vb._lib.check(lib.vb_mvt_objective_f64(ptr(L), ptr(hl), ptr(w), ptr(V), ptr(P), ptr(zu2), ptr(f), ptr(Gm), S, d, 100.0, 2, 2.0, ptr(out[:1]), ptr(out[1:]), ptr(ws), ws.numel(), st()))
torch.cuda.synchronize()
al = lambda n: (n + 255) // 256 * 256
off = 0
regs = {}
for name, n in (('sv', S), ('scal', 4), ('sqrtw', d), ('Q', S * d), ('M', d * d), ('M2', d * d), ('T', d * d), ('Lbar', d * d)):
    regs[name] = ws[off:off + 8 * n].view(torch.float64)
    off += al(8 * n)
for k, v in regs.items():
    print(k, 'finite', fin(v), 'nan', int(torch.isnan(v).sum()), 'absmax', float(v[torch.isfinite(v)].abs().max()) if torch.isfinite(v).any() else None)
print('scal', regs['scal'].tolist(), 'value', float(out[0]))
print('grad mu finite', fin(out[1:1 + d]), 'F part nan', int(torch.isnan(out[1 + d:]).sum()), 'of', out.numel() - 1 - d)
