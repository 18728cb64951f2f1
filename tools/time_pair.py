"""Phase timestamps (clock64) of CTA 0 of the pair kernel + a dump of rb / gm for a small case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import viabel_b200 as vb

def run(N, d, S, show_rb):
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    X = torch.randn(N, d, generator=g, device='cuda', dtype=torch.float64)
    y = torch.where(torch.rand(N, generator=g, device='cuda', dtype=torch.float64) < 0.5, 1.0, -1.0)
    model = vb.LogisticRegression(X, y)
    th = torch.randn(S, d, generator=g, device='cuda', dtype=torch.float64) * 0.05
    bs = torch.randn(S, d, generator=g, device='cuda', dtype=torch.float64).to(torch.float16).to(torch.float64)
    ref = torch.cat(model.sweep(th, bs, None, True))
    model.enable_fast_path()
    grid = vb._lib.lib.vb_device_sm_count()
    dbg = torch.zeros(grid * 49152 + 2048, dtype=torch.float32, device='cuda')
    out = torch.zeros(S + 2 * d, dtype=torch.float64, device='cuda')
    tot = os.environ.get('TOT', '1') == '1'
    model._sweep_fast(th, bs, None, True, out, debug=dbg, ll_total_only=tot)
    dbg.zero_()
    model._sweep_fast(th, bs, None, True, out, debug=dbg, ll_total_only=tot)
    torch.cuda.synchronize()
    if show_rb:
        a = (X * y[:, None]) @ th.T
        R = torch.sigmoid(-a)
        rb_ref = R.sum(dim=1).cpu().numpy()
        D = dbg[:2048].cpu().numpy().reshape(2, 1024)
        for r in range(2):
            print('rank', r, 'rb own rows[:4]', D[r, :4], 'peer rows[:4]', D[r, 128:132], 'gm[:4]', D[r, 256:260], 'ge[:4]', D[r, 384:388])
        print('rb_ref[:4]', rb_ref[:4])
        print('gmu fast[:4]', out[S:S + 4].cpu().numpy(), 'ref', ref[S:S + 4].cpu().numpy())
    else:
        raw = dbg[grid * 49152:].view(torch.int64).cpu().numpy()
        b = raw[5 * 16 + 1]
        print('it5: MMA-thread data-ready time per G1 stage (rel. r_full):', (raw[128:144] - b).tolist())
        print('it5: producer slot-acquire times (rel. r_full):', (raw[192:232] - b).tolist())
        tim = raw[:128].reshape(8, 16)
        t0 = tim[0, 0]
        names = {0: 'iter', 1: 'r_full', 2: 'issued', 8: 'E1start', 6: 'E1loop', 7: 'E1bar', 14: 'E1dsm', 9: 'E1end', 10: 'E2first'}
        acc = {3: 'mma_wait_fill', 4: 'mma_wait_E', 5: 'mma_wait_T', 11: 'e1_math', 12: 'e2_wait_t', 13: 'e2_busy', 15: 'e1_ldtm'}
        for t in range(7):
            print('it', t, ' '.join('%s=%d' % (n, tim[t, i] - t0) for i, n in names.items()), '|', ' '.join('%s=%d' % (n, tim[t, i]) for i, n in acc.items()))

run(128, 128, 64, True)
run(200000, 512, 256, False)
