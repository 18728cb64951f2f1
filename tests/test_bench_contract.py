"""CPU: the bench contract.  The reference arm (numpy oracle on host cores) runs here on a tiny shape and must
print ONE JSON line with the keys the driver reads; the keys of the GPU arm's line are checked on the committed
round-1 lines under profiles/."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e'}


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--n-obs', '4000',
                          '--dim', '16', '--mc', '8', '--steps', '2', '--warmup', '1'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and BASE_KEYS <= set(d)
    assert d['metric'] == 'elbo_grad_iters_per_sec' and d['unit'] == 'iter/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 2 and d['warmup'] == 1
    assert set(d['e2e']) == {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'}
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                          '--n-obs', '4000', '--dim', '16', '--mc', '8', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_committed_gpu_lines_have_the_contract_keys():
    for name, gpus in (('bench_r01_1gpu.json', 1), ('bench_r01_8gpu.json', 8)):
        with open(os.path.join(ROOT, 'profiles', name)) as f:
            d = json.loads(f.read())
        assert BASE_KEYS | {'clocks', 'gpu_launches', 'roofline', 'cpu_baseline'} <= set(d)
        assert d['n_gpus'] == gpus and d['gpu_launches'] > 0 and d['warmup'] >= 3
        assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(d['roofline'])
        assert abs(d['roofline']['frac'] - d['roofline']['achieved'] / d['roofline']['peak']) < 1e-9
        assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(d['clocks'])
        assert not {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'} & set(d['clocks']['reasons'])
        if gpus == 1:
            assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(d['cpu_baseline'])
            assert d['e2e']['value'] != d['value'] and d['e2e']['h2d_bytes_per_step'] > 0
