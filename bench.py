#!/usr/bin/env python
"""bench.py -- ELBO-gradient iterations/s on BASELINE.json configs[1]:
Bayesian logistic regression N=1e6, d=512, S=256, MFGaussian + RMSProp (synthetic data).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--path f64|fast] [--impl reference]

One "step" = objective(var_param) (Philox draws -> fused sweep over all N observations ->
value + gradient) followed by the fused RMSProp update, exactly the three hot-path calls of the
reference loop (optimization.py:95-98).  For N > 1 the observations are sharded over the ranks
(strong scaling of the same problem) and the S + 2d partial sums are all-reduced with NCCL.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS, DIM, S_MC = 1000000, 512, 256
DATA_SEED, DRAW_SEED = 20260117, 1234
CPU_SAMPLE_ROWS = 100000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--path', default=os.environ.get('VB_BENCH_PATH', 'auto'), choices=['auto', 'f64', 'fast'])
    ap.add_argument('--n-obs', type=int, default=N_OBS)
    ap.add_argument('--dim', type=int, default=DIM)
    ap.add_argument('--mc', type=int, default=S_MC)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-psis', action='store_true')
    ap.add_argument('--psis-draws', type=int, default=100000000)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU arm: the numpy oracle (a port of the reference iteration; the reference's own autograd
# path cannot run in this image -- autograd/paragami are absent) on a bounded sample of rows.
# ----------------------------------------------------------------------------------------------
def host_problem(n_rows, d, seed):
    rs = np.random.RandomState(seed)
    X = rs.standard_normal((n_rows, d))
    beta = rs.standard_normal(d) / np.sqrt(d)
    y = np.where(rs.random_sample(n_rows) < 1.0 / (1.0 + np.exp(-(X @ beta))), 1.0, -1.0)
    return X, y


def cpu_iterations(n_rows, d, S, steps, warmup):
    """Returns seconds per iteration of the oracle on n_rows observations."""
    from oracle import viabel_oracle as vo
    X, y = host_problem(n_rows, d, DATA_SEED)
    rs = np.random.RandomState(DRAW_SEED)
    vp = vo.mfg_init_param(d)
    state = {}
    times = []
    for it in range(warmup + steps):
        eps = rs.standard_normal((S, d))
        t0 = time.perf_counter()
        vp, _, _ = vo.elbo_step_logistic(vp, eps, X, y, state, lr=0.01)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas']
        return max(n) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rows = min(CPU_SAMPLE_ROWS, args.n_obs)
    sec = cpu_iterations(rows, args.dim, args.mc, args.steps, max(1, min(args.warmup, 1)))
    scaled = sec * (args.n_obs / rows)          # seconds per full-size iteration
    value = 1.0 / scaled
    sample = ('oracle (numpy float64 port of the reference iteration) on %d of %d rows, %d steps; '
              'time scaled linearly in rows' % (rows, args.n_obs, args.steps))
    line = {
        'impl': 'reference', 'metric': 'elbo_grad_iters_per_sec', 'value': value, 'unit': 'iter/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': scaled * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'bayes-logistic N=%d d=%d S=%d MFGaussian+RMSProp (BASELINE configs[1])' % (args.n_obs, args.dim, args.mc)},
        'cpu_baseline': {'value': value, 'unit': 'iter/s', 'cores': cpu_threads(), 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'iter/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out = {'sm_mhz': float(np.median(busy)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                   'samples': len(sm), 'power_w_max': max(power)}
        return out


def profiled_traffic(*kernels):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernels, from the committed
    ncu --set full captures (profiles/traffic_r01.json); None when a kernel has not been captured."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'traffic_r01.json')) as f:
            t = json.load(f)
        return float(sum(t[k] for k in kernels))
    except (OSError, KeyError, ValueError):
        return None


def measured_peak(name, fallback):
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)[name]), 'measured'
    except Exception:
        return fallback, 'fallback'


def matmul_peak_tflops(torch, dtype, tf32, n):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device='cuda', dtype=dtype)
    b = torch.randn(n, n, device='cuda', dtype=dtype)
    best = 1e9
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = min(best, e0.elapsed_time(e1) * 1e-3)
    torch.backends.cuda.matmul.allow_tf32 = False
    return 2.0 * n ** 3 / best / 1e12


def bench_psis(torch, vb, args):
    """Second headline metric (BASELINE.json: 'PSIS draws/s'): psislw + CUBO/ELBO moments on n
    float64 log-weights resident in HBM (BASELINE configs[4] size, one GPU).  HBM roofline with
    24 algorithmic bytes per draw (lw read twice, smoothed weights written once)."""
    n = args.psis_draws
    gen = torch.Generator(device='cuda')
    gen.manual_seed(DATA_SEED + 5)
    # log p - log q for p = t_10, q = t_40 per coordinate summed over 4 coordinates (heavy-ish tail)
    lw = torch.zeros(n, device='cuda', dtype=torch.float64)
    for _ in range(2):
        z = torch.randn(n, generator=gen, device='cuda', dtype=torch.float64)
        lw += -5.5 * torch.log1p(z * z / 10.0) + 20.5 * torch.log1p(z * z / 40.0)
        del z
    out = torch.empty_like(lw)
    for _ in range(3):
        vb.psislw_device(lw, out)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _, res, _, _ = vb.psislw_device(lw, out)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    r = res.cpu().numpy()
    # diagnostics-only mode (k-hat and the CUBO / ELBO moments, no smoothed output): 16 algorithmic bytes per draw
    for _ in range(2):
        vb.psislw_device(lw, None)
    e0.record()
    for _ in range(reps):
        vb.psislw_device(lw, None)
    e1.record()
    torch.cuda.synchronize()
    sec_diag = e0.elapsed_time(e1) * 1e-3 / reps
    hbm, how = measured_peak('hbm_gbs', 6650.0)
    achieved = 24.0 * n / sec / 1e9
    # end to end from HOST memory: H2D of the weights, PSIS, D2H of k-hat and the smoothed weights
    host = torch.empty(min(n, 20000000), dtype=torch.float64).pin_memory()
    host.copy_(lw[:host.numel()])
    hout = torch.empty_like(host).pin_memory()
    dev_in = torch.empty(host.numel(), device='cuda', dtype=torch.float64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_in.copy_(host, non_blocking=True)
    _, res2, _, _ = vb.psislw_device(dev_in, dev_in)
    hout.copy_(dev_in, non_blocking=True)
    k2 = float(res2[0].item())
    torch.cuda.synchronize()
    e2e = host.numel() / (time.perf_counter() - t0)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import viabel_oracle as vo
        m = min(n, 5000000)
        sample = lw[:m].cpu().numpy()
        t0 = time.perf_counter()
        with np.errstate(all='ignore'):
            o, k = vo.psislw_1d(sample)
            vo.divergence_bound(o)
        cpu = {'value': m / (time.perf_counter() - t0), 'unit': 'draws/s', 'cores': 1, 'kind': 'port',
               'sample': 'numpy oracle psislw + divergence_bound on the first %d draws' % m}
    del lw, out
    return {'metric': 'psis_draws_per_sec', 'value': n / sec, 'unit': 'draws/s', 'n_draws': n, 'ms': sec * 1e3,
            'khat': float(r[0]), 'n_tail': int(r[2]), 'status': int(r[6]),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm, 'unit': 'GB/s', 'frac': achieved / hbm,
                         'traffic': profiled_traffic('psis_pass_a_kernel', 'psis_pass_b_kernel') if n == 100000000 else None,
                         'traffic_note': 'the two streaming passes (the short kernels between them touch < 10 MB)',
                         'peak_source': how, 'algorithmic_bytes_per_draw': 24},
            'diagnostics_only': {'value': n / sec_diag, 'unit': 'draws/s', 'ms': sec_diag * 1e3,
                                 'algorithmic_bytes_per_draw': 16, 'achieved_gbs': 16.0 * n / sec_diag / 1e9,
                                 'frac': 16.0 * n / sec_diag / 1e9 / hbm},
            'e2e': {'value': e2e, 'unit': 'draws/s', 'n_draws': host.numel(), 'h2d_bytes': host.numel() * 8,
                    'd2h_bytes': host.numel() * 8 + 8, 'khat': k2},
            'cpu_baseline': cpu}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import viabel_b200 as vb
    from viabel_b200.parallel import shard_rows

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    N, d, S = args.n_obs, args.dim, args.mc
    # this rank's rows [lo, hi) of the N x d problem; rank r seeds its rows with DATA_SEED + r
    lo, hi = shard_rows(N, rank, world)
    gen = torch.Generator(device=dev)
    gen.manual_seed(DATA_SEED + 7919)
    beta = torch.randn(d, generator=gen, device=dev, dtype=torch.float64) / np.sqrt(d)
    gen.manual_seed(DATA_SEED + rank)
    X = torch.randn(hi - lo, d, generator=gen, device=dev, dtype=torch.float64)
    p = torch.sigmoid(X @ beta)
    y = torch.where(torch.rand(hi - lo, generator=gen, device=dev, dtype=torch.float64) < p, 1.0, -1.0)
    del p

    path = 'fast' if args.path == 'auto' else args.path
    model = vb.LogisticRegression(X, y, prior_scale=10.0, sharded=world > 1)
    approx = vb.MFGaussian(d, seed=DRAW_SEED)
    if path == 'fast':
        model.enable_fast_path()          # tcgen05 + TMA, fp16 hi/lo operand splits, 1e-4 tolerance
        approx.quantize_draws = 2         # fp16-exact Philox normals (exact tensor-core operands)
    objective = vb.ExclusiveKL(approx, model, S)
    opt = vb.RMSProp(0.01)
    vp = torch.as_tensor(approx.init_param(), device=dev)
    # own kernels per step.  fast: philox, sample, operand pack, pair sweep, partial reduction, value, grad, rmsprop;
    # f64: philox, sample, pack, sweep, 3x reduce, value, grad, rmsprop.  (The NCCL all-reduce at N > 1 is not ours.)
    launches_per_step = 8 if path == 'fast' else 10

    def step():
        value, grad = objective(vp)
        opt._fused_step(vp, grad, False)
        return value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    elapsed = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    # nvidia-smi samples every 100 ms; a short timed region (K steps of ~1 ms) can fall between two samples,
    # so the same step keeps running (untimed; the same count on every rank, the step contains the all-reduce)
    # until the sampler has seen the GPU under this load
    n_extra = 0 if elapsed >= 0.6 else min(5000, int(0.6 / max(elapsed / args.steps, 1e-5)))
    for _ in range(n_extra):
        step()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_per_step = elapsed / args.steps * 1e3

    # ---- end-to-end through the public API with HOST buffers (numpy in, numpy out) -----------
    vp_host = approx.init_param()
    opt2 = vb.RMSProp(0.01)
    for _ in range(2):
        v, g = objective(vp_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, g = objective(vp_host)                   # H2D var_param, D2H value + gradient
        vp_host = vp_host - 0.01 * opt2.descent_direction(g)
    barrier()
    e2e_sec = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())

    # ---- dominant kernel alone (the fused sweep), CUDA events on the launching stream --------
    theta = approx.sample(vp, S)
    base = approx.last_base
    for _ in range(2):
        model.sweep(theta, base, None, True)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, min(args.steps, 10))
    k0.record()
    for _ in range(reps):
        model.sweep(theta, base, None, True)
    k1.record()
    torch.cuda.synchronize()
    sweep_sec = k0.elapsed_time(k1) * 1e-3 / reps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops = 4.0 * (hi - lo) * d * S                  # algorithmic flops of one sweep on this rank
    if path == 'f64':
        peak = matmul_peak_tflops(torch, torch.float64, False, 4096)
        peak_note = 'cuBLAS fp64 matmul 4096^3 measured in this run'
    else:
        peak = matmul_peak_tflops(torch, torch.float32, True, 8192)
        peak_note = 'cuBLAS tf32 matmul 8192^3 measured in this run'
    achieved = flops / sweep_sec / 1e12
    bf16_peak, how = measured_peak('bf16_tflops', 1590.0)
    traffic = None
    if path == 'fast' and world == 1 and (N, d, S) == (1000000, 512, 256):
        traffic = profiled_traffic('glm_fast_pair_kernel')       # ncu --set full capture of this very launch shape
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                'frac': achieved / peak, 'traffic': traffic, 'kernel': 'glm_sweep_%s' % path,
                'kernel_ms': sweep_sec * 1e3, 'peak_note': peak_note,
                'bf16_peak_tflops': bf16_peak, 'bf16_peak_source': how,
                'algorithmic_flops_per_launch': flops, 'algorithmic_bytes_per_launch': (hi - lo) * d * 8.0}

    psis = bench_psis(torch, vb, args) if world == 1 and not args.no_psis else None

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        rows = min(CPU_SAMPLE_ROWS, N)
        sec = cpu_iterations(rows, d, S, 3, 1) * (N / rows)
        cpu_baseline = {'value': 1.0 / sec, 'unit': 'iter/s', 'cores': cpu_threads(), 'kind': 'port',
                        'sample': 'numpy float64 oracle, 3 iterations on %d of %d rows, scaled linearly' % (rows, N)}

    line = {
        'metric': 'elbo_grad_iters_per_sec', 'value': 1e3 / ms_per_step, 'unit': 'iter/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64' if path == 'f64' else 'f16x2-split/f32',
        'data': 'synthetic',
        'config': {'workload': 'bayes-logistic N=%d d=%d S=%d MFGaussian+RMSProp (BASELINE configs[%d])'
                   % (N, d, S, 2 if (N, d) == (10000000, 1024) else 1),
                   'path': path, 'rows_per_rank': hi - lo, 'l2': 'inputs larger than L2 (X = %.2f GB per rank)'
                   % ((hi - lo) * d * 8 / 1e9)},
        'clocks': clocks,
        'e2e': {'value': 1.0 / e2e_sec, 'unit': 'iter/s', 'h2d_bytes_per_step': 2 * d * 8,
                'd2h_bytes_per_step': (1 + 2 * d) * 8},
        'gpu_launches': launches_per_step * args.steps,
        'roofline': roofline,
        'cpu_baseline': cpu_baseline,
        'psis': psis,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
