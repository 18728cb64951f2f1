"""GPU parity for the fused three-kernel step (engine.cu / viabel_b200.engine.FusedStep): the whole
reference iteration objective(var_param) -> descent_direction -> update (optimization.py:95-98) against the
numpy oracle, through the C ABI, with injected draws; graph replay against eager enqueue; the draw stream
against vb_philox_*; the peer-memory communicator and the in-kernel exchange with two ranks driven by one
process (two streams on one GPU)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import relerr
from _problems import logistic_problem, target_params

pytestmark = pytest.mark.gpu
TOL64 = 1e-10
TOL_FAST = 1e-4


@pytest.fixture(scope='module')
def vb():
    import viabel_b200
    return viabel_b200


@pytest.fixture(scope='module')
def vo():
    from oracle import viabel_oracle
    return viabel_oracle


def _oracle_loop(vo, vp, bases, X, y, opt, family='gaussian', df=None, path=False, lr=0.01):
    state = {}
    out = []
    for base in bases:
        v, g, _ = vo.exclusive_kl_meanfield(vp, base, lambda th: vo.logistic_logp_grad(th, X, y, 10.0),
                                            family, df, path)
        d = vo.rmsprop_direction(state, g) if opt == 'rmsprop' else vo.adam_direction(state, g)
        vp = vp - lr * d
        out.append((v, g, vp.copy()))
    return out


@pytest.mark.parametrize('N,d,S', [(1003, 13, 7), (5000, 64, 32), (2048, 300, 256), (700, 512, 100)])
@pytest.mark.parametrize('fam', ['gaussian', 'student'])
@pytest.mark.parametrize('path', [False, True])
def test_fused_step_f64_vs_oracle(vb, vo, N, d, S, fam, path):
    """Float64 sweep inside the fused step: value, gradient and the updated parameter over 3 iterations."""
    from viabel_b200.engine import FusedStep
    X, y, beta = logistic_problem(N, d, seed=N + d)
    rs = np.random.RandomState(S)
    model = vb.LogisticRegression(X, y, prior_scale=10.0)
    approx = vb.MFGaussian(d) if fam == 'gaussian' else vb.MFStudentT(d, 7.5)
    df = None if fam == 'gaussian' else 7.5
    obj = vb.ExclusiveKL(approx, model, S, use_path_deriv=path)
    for opt_name, opt in (('rmsprop', vb.RMSProp(0.01)), ('adam', vb.Adam(0.01))):
        bases = [rs.randn(S, d) if fam == 'gaussian' else rs.standard_t(7.5, size=(S, d)) for _ in range(3)]
        vp0 = np.concatenate([0.5 * beta, -1.0 + 0.1 * rs.randn(d)])
        ref = _oracle_loop(vo, vp0, bases, X, y, opt_name, fam, df, path)
        eng = FusedStep(obj, opt, ring=4, hist_len=3, inject_base=True, want_grad_hist=True)
        eng.set_param(vp0)
        for k, base in enumerate(bases):
            eng.base.copy_(torch.as_tensor(base, device='cuda'))
            eng.run(1, use_graph=(k > 0))          # eager first, graph replay afterwards
            v0, g0, vp_ref = ref[k]
            assert relerr(float(eng.value[0]), v0) < TOL64, (opt_name, k, 'value')
            assert relerr(eng.grad.cpu().numpy(), g0) < TOL64, (opt_name, k, 'grad')
            assert relerr(eng.vp.cpu().numpy(), vp_ref) < TOL64, (opt_name, k, 'param')
        assert relerr(eng.value_hist.cpu().numpy(), [r[0] for r in ref]) < TOL64
        assert relerr(eng.last_rows(eng.param_hist, 3).cpu().numpy(), np.stack([r[2] for r in ref])) < TOL64
        assert relerr(eng.last_rows(eng.grad_hist, 2).cpu().numpy(), np.stack([r[1] for r in ref[1:]])) < TOL64


@pytest.mark.parametrize('N,d,S', [(1003, 13, 7), (20000, 512, 256), (4100, 130, 100), (257, 1024, 256)])
def test_fused_step_fast_vs_oracle(vb, vo, N, d, S):
    """Tensor-core sweep inside the fused step (fp16-exact injected draws), 1e-4 on value / gradient / update."""
    from viabel_b200.engine import FusedStep
    X, y, beta = logistic_problem(N, d, seed=N + d)
    rs = np.random.RandomState(S)
    model = vb.LogisticRegression(X, y, prior_scale=10.0).enable_fast_path()
    approx = vb.MFGaussian(d)
    obj = vb.ExclusiveKL(approx, model, S)
    bases = [rs.randn(S, d).astype(np.float16).astype(np.float64) for _ in range(3)]
    for vp0 in (vo.mfg_init_param(d), np.concatenate([beta + 0.01 * rs.randn(d), -3.5 + 0.1 * rs.randn(d)])):
        ref = _oracle_loop(vo, vp0, bases, X, y, 'rmsprop')
        eng = FusedStep(obj, vb.RMSProp(0.01), inject_base=True)
        eng.set_param(vp0)
        for k, base in enumerate(bases):
            eng.base.copy_(torch.as_tensor(base, device='cuda'))
            eng.run(1)
            v0, g0, vp_ref = ref[k]
            assert relerr(float(eng.value[0]), v0) < TOL_FAST
            assert relerr(eng.grad.cpu().numpy(), g0) < TOL_FAST
            # the RMSProp direction is g / sqrt(nu): a relative gradient error passes through unchanged
            assert relerr(eng.vp.cpu().numpy() - vp0, vp_ref - vp0) < 2 * TOL_FAST


def test_fused_draws_match_the_philox_stream(vb):
    """The draws generated inside the step's first kernel are the family's Philox stream, element for element,
    for both families, quantised or not, and the stream position carries across calls and graph replays."""
    N, d, S = 3000, 70, 33                                  # odd S*d: exercises the even-rounded stride
    X, y, beta = logistic_problem(N, d, seed=5)
    model = vb.LogisticRegression(X, y)
    cases = ((lambda: vb.MFGaussian(d, seed=77), 0), (lambda: vb.MFGaussian(d, seed=78), 2),
             (lambda: vb.MFStudentT(d, 6.0, seed=79), 0), (lambda: vb.MFStudentT(d, 6.0, seed=80), 2))
    for make, q in cases:
        approx, twin = make(), make()
        approx.quantize_draws = twin.quantize_draws = q
        obj = vb.ExclusiveKL(approx, model, S)
        vp = torch.as_tensor(np.concatenate([beta, -2.0 * np.ones(d)]), device='cuda')
        for _ in range(3):
            obj(vp)                                         # fused path (device tensor in)
            expect = twin.base_draws(S)
            assert torch.equal(approx.last_base, expect)
        v_host, g_host = obj(vp.cpu().numpy())              # host path: one graph with the two copies
        assert torch.equal(approx.last_base, twin.base_draws(S))
        assert approx._offset == twin._offset
        # the unfused kernels on the same draws give the same numbers
        from viabel_b200 import objectives
        v2, g2 = objectives._mf_objective(approx, model, S, 0, 0.0, vp, base=approx.last_base.clone())
        assert relerr(v_host, float(v2)) < 1e-13 and relerr(g_host, g2.cpu().numpy()) < 1e-12


def test_fused_optimize_matches_unfused_loop(vb, vo):
    """RMSProp.optimize() through the graph-replayed engine == the oracle loop on the same Philox draws
    (read back from a twin family), including the iterate-average window of optimization.py:103-121."""
    N, d, S, iters = 4000, 24, 16, 37
    X, y, beta = logistic_problem(N, d, seed=12)
    model = vb.LogisticRegression(X, y)
    approx = vb.MFGaussian(d, seed=5)
    twin = vb.MFGaussian(d, seed=5)
    opt = vb.RMSProp(0.05)
    opt.progress = False
    res = opt.optimize(iters, vb.ExclusiveKL(approx, model, S), approx.init_param())
    bases = [twin.base_draws(S).cpu().numpy() for _ in range(iters)]
    ref = _oracle_loop(vo, approx.init_param(), bases, X, y, 'rmsprop', lr=0.05)
    assert relerr(res['value_history'], [r[0] for r in ref]) < 1e-9
    hist = np.stack([r[2] for r in ref])
    k = iters - 1
    window = max(1, int(k * 0.2))
    assert relerr(res['opt_param'], hist[-window:].mean(axis=0)) < 1e-9
    assert relerr(res['variational_param_history'], hist[-res['variational_param_history'].shape[0]:]) < 1e-9
    # the optimiser state persists across optimize() calls, like the reference's
    res2 = opt.optimize(3, vb.ExclusiveKL(approx, model, S), hist[-1])
    bases2 = [twin.base_draws(S).cpu().numpy() for _ in range(3)]
    state = {}
    vp = approx.init_param()
    for b in bases:
        _, g, _ = vo.exclusive_kl_meanfield(vp, b, lambda th: vo.logistic_logp_grad(th, X, y, 10.0))
        vp = vp - 0.05 * vo.rmsprop_direction(state, g)
    vals = []
    for b in bases2:
        v, g, _ = vo.exclusive_kl_meanfield(vp, b, lambda th: vo.logistic_logp_grad(th, X, y, 10.0))
        vp = vp - 0.05 * vo.rmsprop_direction(state, g)
        vals.append(v)
    assert relerr(res2['value_history'], vals) < 1e-9


def test_graph_replay_is_bitwise_eager(vb):
    from viabel_b200.engine import FusedStep
    N, d, S = 9000, 128, 64
    X, y, beta = logistic_problem(N, d, seed=3)
    model = vb.LogisticRegression(X, y).enable_fast_path()
    outs = []
    for use_graph in (False, True):
        approx = vb.MFGaussian(d, seed=9)
        approx.quantize_draws = 2
        eng = FusedStep(vb.ExclusiveKL(approx, model, S), vb.Adam(0.02), ring=8, hist_len=21)
        eng.set_param(approx.init_param())
        eng.run(21, use_graph=use_graph)
        outs.append((eng.vp.clone(), eng.value_hist.clone(), approx._offset))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]


class _LocalComm(object):
    """One rank of a communicator whose peers live in this process (vb_comm_connect_ptrs)."""

    def __init__(self, lib, rank, world, slot_bytes):
        self.handle = ctypes.c_void_p()
        hbuf = (ctypes.c_ubyte * 64)()
        assert lib.vb_comm_create(ctypes.byref(self.handle), rank, world, slot_bytes, hbuf) == 0

    @staticmethod
    def connect(lib, comms):
        ptrs = (ctypes.c_void_p * len(comms))(*[lib.vb_comm_buffer(c.handle) for c in comms])
        for c in comms:
            assert lib.vb_comm_connect_ptrs(c.handle, ptrs) == 0


def test_comm_allreduce_two_ranks_one_process(vb):
    """vb_comm_allreduce_sum_f64 with 3 ranks driven from one process on 3 streams: rank-ordered sums,
    identical on every rank, repeated (parity double-buffering) without resetting anything."""
    lib = vb._lib.lib
    world, n = 3, 1500
    comms = [_LocalComm(lib, r, world, n * 8) for r in range(world)]
    _LocalComm.connect(lib, comms)
    streams = [torch.cuda.Stream() for _ in range(world)]
    rs = np.random.RandomState(0)
    for rep in range(5):
        host = [rs.randn(n) for _ in range(world)]
        bufs = [torch.as_tensor(h, device='cuda') for h in host]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                assert lib.vb_comm_allreduce_sum_f64(comms[r].handle, bufs[r].data_ptr(), n, streams[r].cuda_stream) == 0
        torch.cuda.synchronize()
        expect = host[0] + host[1] + host[2]                 # rank order
        for r in range(world):
            assert lib.vb_comm_error(comms[r].handle) == 0
            assert np.array_equal(bufs[r].cpu().numpy(), expect), (rep, r)
    for c in comms:
        lib.vb_comm_destroy(c.handle)


def test_fused_step_two_ranks_one_process(vb, vo):
    """The in-kernel exchange of the fused step: the observations split over two 'ranks' (two engines on two
    streams of one GPU, peer buffers connected by pointer) give the single-rank result, and both ranks hold
    bit-identical parameters after every step."""
    from viabel_b200.engine import FusedStep
    lib = vb._lib.lib
    N, d, S = 1500, 40, 24                                   # few tiles: both sweeps fit on the GPU together
    X, y, beta = logistic_problem(N, d, seed=21)
    cut = 700
    rs = np.random.RandomState(2)
    bases = [rs.randn(S, d) for _ in range(4)]
    vp0 = np.concatenate([0.3 * beta, -1.5 * np.ones(d)])
    ref = _oracle_loop(vo, vp0, bases, X, y, 'rmsprop')
    for fast in (False, True):
        comms = [_LocalComm(lib, r, 2, (S + 2 * d) * 8) for r in range(2)]
        _LocalComm.connect(lib, comms)
        engs, streams = [], [torch.cuda.Stream(), torch.cuda.Stream()]
        for r, (lo, hi) in enumerate(((0, cut), (cut, N))):
            model = vb.LogisticRegression(X[lo:hi], y[lo:hi])
            if fast:
                model.enable_fast_path()
            eng = FusedStep(vb.ExclusiveKL(vb.MFGaussian(d), model, S), vb.RMSProp(0.01), inject_base=True)
            eng.comm = comms[r]
            eng.set_param(vp0)
            engs.append(eng)
        tol = TOL_FAST if fast else TOL64
        for k, base in enumerate(bases):
            b = torch.as_tensor(base, device='cuda')
            if fast:
                b = b.to(torch.float16).to(torch.float64)
            for eng in engs:
                eng.base.copy_(b)
            torch.cuda.synchronize()
            for eng, st in zip(engs, streams):
                with torch.cuda.stream(st):
                    eng.enqueue()
            torch.cuda.synchronize()
            for c in comms:
                assert lib.vb_comm_error(c.handle) == 0
            assert torch.equal(engs[0].vp, engs[1].vp) and torch.equal(engs[0].grad, engs[1].grad)
            if not fast:
                assert relerr(float(engs[0].value[0]), ref[k][0]) < tol
                assert relerr(engs[0].grad.cpu().numpy(), ref[k][1]) < tol
                assert relerr(engs[0].vp.cpu().numpy(), ref[k][2]) < tol
        if fast:
            refq = _oracle_loop(vo, vp0, [np.asarray(b, dtype=np.float16).astype(np.float64) for b in bases], X, y, 'rmsprop')
            assert relerr(engs[0].vp.cpu().numpy() - vp0, refq[-1][2] - vp0) < 4 * tol
        del engs
        for c in comms:
            lib.vb_comm_destroy(c.handle)


def _ar1_chains(n, P, seed):
    rs = np.random.RandomState(seed)
    phi = rs.uniform(0.0, 0.95, size=P)
    x = np.zeros((n, P))
    e = rs.randn(n, P)
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    x += np.linspace(0, 1.5, n)[:, None] * (rs.rand(P) < 0.2)          # some coordinates still drifting
    x[:, 3] = 0.25                                                     # a constant coordinate
    return x


@pytest.mark.parametrize('n,P,ring,shift', [(900, 37, 900, 0), (1500, 130, 2000, 700), (640, 1024, 640, 411)])
def test_ring_statistics_vs_host_mirror(vb, n, P, ring, shift):
    """csrc/faso.cu on a (wrapped) device ring against the numpy mirror of _mc_diagnostics.py (itself pinned to the
    reference in tests/golden/mc_diagnostics.npz): split-R-hat per window, window means, ESS and MCSE."""
    from viabel_b200 import _mc_diagnostics as mc
    x = _ar1_chains(n, P, seed=n + P)
    hist = torch.zeros(ring, P, dtype=torch.float64, device='cuda')
    idx = (np.arange(n) + shift) % ring                                # row t of the chain lives at (t + shift) % ring
    hist[torch.as_tensor(idx, device='cuda')] = torch.as_tensor(x, device='cuda')
    end = (shift + n) % ring
    stats = mc.RingStats(hist, ring, P)
    windows = np.linspace(200, int(0.95 * n), num=5, dtype=int)
    got = stats.rhat_max(end, windows)
    with np.errstate(all='ignore'):
        want = np.array([np.max(mc.compute_R_hat(x[-int(w):], 0)) for w in windows])
    np.testing.assert_allclose(got, want, rtol=1e-10)
    ok, best = stats.convergence_check(end, windows)
    ok0, best0 = mc.R_hat_convergence_check(x, windows)
    assert ok == ok0 and best == best0
    for W in (201, 400, n):
        mean = stats.window_mean(end, W).cpu().numpy()
        np.testing.assert_allclose(mean, x[-W:].mean(axis=0), rtol=1e-12, atol=1e-13)
        ess, mcse, mean2 = stats.mcse(end, W)
        with np.errstate(all='ignore'):
            ess0, mcse0 = mc.MCSE(x[-W:])
        ess0 = np.asarray(ess0)
        fin = np.isfinite(ess0)
        assert np.array_equal(np.isnan(ess), np.isnan(ess0))
        np.testing.assert_allclose(ess[fin], ess0[fin], rtol=1e-7)
        np.testing.assert_allclose(mcse[fin], mcse0[fin], rtol=1e-7)


def test_faso_fused_matches_unfused(vb, monkeypatch, capsys):
    """FASO (optimization.py:521-633) through graph-replayed steps + device-ring statistics takes the same decisions
    (k_Rhat, k_conv, k_stopped) and returns the same iterate average as the per-iteration loop on the same draws."""
    import viabel_b200.engine as engine
    N, d, S = 3000, 6, 12
    X, y, beta = logistic_problem(N, d, seed=44)
    model = vb.LogisticRegression(X, y)
    out = {}
    for fused in (True, False):
        if not fused:
            monkeypatch.setattr(engine, 'fused_step_supported', lambda *a, **k: False)
        approx = vb.MFGaussian(d, seed=17)
        sgo = vb.RMSProp(0.05, diagnostics=True)
        sgo.progress = False
        faso = vb.FASO(sgo, W_min=100, k_check=100, mcse_threshold=0.2)
        out[fused] = faso.optimize(3000, vb.ExclusiveKL(approx, model, S), approx.init_param())
    a, b = out[True], out[False]
    # the R-hat decisions are deterministic; where the MCSE re-checks fall afterwards depends on the measured
    # optimisation / MCSE time ratio (optimization.py:599-605), so k_stopped may differ between the two loops
    assert a['k_Rhat'] == b['k_Rhat'] and a['k_conv'] == b['k_conv'] and a['k_conv'] is not None
    assert a['k_stopped'] is not None and b['k_stopped'] is not None
    n = min(len(a['value_history']), len(b['value_history']))
    assert a['variational_param_history'].shape[1:] == b['variational_param_history'].shape[1:]
    assert relerr(a['value_history'][:n], b['value_history'][:n]) < 1e-9
    assert relerr(a['variational_param_history'][:n], b['variational_param_history'][:n]) < 1e-9
    assert relerr(a['grad_history'][:n], b['grad_history'][:n]) < 1e-8
    assert relerr(a['descent_dir_history'][:n], b['descent_dir_history'][:n]) < 1e-8
    # the first MCSE check happens at the same iteration over the same window in both loops
    assert a['ess_and_mcse_k_history'][0] == b['ess_and_mcse_k_history'][0]
    assert relerr(a['mcse_history'][0], b['mcse_history'][0]) < 1e-6
    assert relerr(a['ess_history'][0], b['ess_history'][0]) < 1e-6
    assert relerr(a['iterate_average_history'][1], b['iterate_average_history'][1]) < 1e-9


@pytest.mark.parametrize('family', ['mfg', 'mft'])
def test_faso_device_ring_matches_host_lists(vb, family):
    """The general (unfused) FASO loop keeps its iterate history in a device ring and runs split-R-hat / ESS / MCSE on
    it (csrc/faso.cu); with the ring disabled it falls back to the reference's host lists and numpy statistics
    (optimization.py:546-605, _mc_diagnostics.py).  Same draws -> same decisions and the same iterate average."""
    d = 5
    mean, sd = target_params(d, seed=23)
    model = vb.GaussianTarget(mean, sd)
    out = {}
    for ring in (True, False):
        approx = vb.MFGaussian(d, seed=29) if family == 'mfg' else vb.MFStudentT(d, 9, seed=29)
        sgo = vb.RMSProp(0.05, diagnostics=True)
        sgo.progress = False
        faso = vb.FASO(sgo, W_min=100, k_check=100, mcse_threshold=0.2)
        if not ring:
            faso._device_ring_bytes = 0
        out[ring] = faso.optimize(2500, vb.ExclusiveKL(approx, model, 10), approx.init_param())
    a, b = out[True], out[False]
    assert a['k_Rhat'] == b['k_Rhat'] and a['k_conv'] == b['k_conv'] and a['k_conv'] is not None
    n = min(len(a['value_history']), len(b['value_history']))
    assert relerr(a['variational_param_history'][:n], b['variational_param_history'][:n]) < 1e-12
    assert a['ess_and_mcse_k_history'][0] == b['ess_and_mcse_k_history'][0]
    assert relerr(a['mcse_history'][0], b['mcse_history'][0]) < 1e-6
    assert relerr(a['ess_history'][0], b['ess_history'][0]) < 1e-6
    assert relerr(a['iterate_average_history'][1], b['iterate_average_history'][1]) < 1e-10
