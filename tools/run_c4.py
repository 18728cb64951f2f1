"""BASELINE configs[3]: hierarchical linear regression d=2048, MultivariateT + AlphaDivergence(2), S=256.
Times one objective evaluation + RMSProp step on one GPU (the d x d algebra is replicated)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import viabel_b200 as vb

G, p, n_per = 65, 31, 200
rs = np.random.RandomState(20260118)
N = G * n_per
group = np.repeat(np.arange(G), n_per)
X = rs.randn(N, p)
m = rs.randn(p)
beta = m + 0.5 * rs.randn(G, p)
y = np.sum(X * beta[group], axis=1) + 0.3 * rs.randn(N)
model = vb.HierarchicalLinearRegression(X, y, group, G)
d = model.dim
approx = vb.MultivariateT(d, 100)
obj = vb.AlphaDivergence(approx, model, 256, 2.0)
vp = torch.as_tensor(approx.init_param(), device='cuda')
vp[d:] = torch.as_tensor(np.log(0.05) * (approx.init_param()[d:] != 0), device='cuda')     # Sigma = 0.0025 I
opt = vb.RMSProp(1e-6)   # RMSProp first steps are sign(g)*lr on all 2.1M coordinates: keep L near its start
for it in range(3):
    v, g = obj(vp)
    print('warm', it, float(v), float(g.abs().max()), bool(torch.isfinite(g).all()), float(vp.abs().max()), file=sys.stderr)
    opt._fused_step(vp, g, False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    v, g = obj(vp); opt._fused_step(vp, g, False)
e1.record(); torch.cuda.synchronize()
print(json.dumps({'config': 'hier-linear d=%d N=%d MultivariateT(df=100) AlphaDivergence(2) S=256' % (d, N),
                  'var_param_dim': int(vp.numel()), 'ms_per_iter': e0.elapsed_time(e1) / K, 'value': float(v),
                  'grad_finite': bool(torch.isfinite(g).all())}))
