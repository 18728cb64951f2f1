"""Build libviabel_b200.so in-tree with nvcc for sm_100a (no JIT cache, no pip).

    python viabel_b200/csrc/build.py            # incremental
    python viabel_b200/csrc/build.py --force

Run as a script (or through __graft_entry__.build()); it must not import the package,
whose __init__ loads the library being built.
"""
import os
import subprocess
import sys

CSRC = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.dirname(CSRC)
LIB = os.path.join(HERE, 'libviabel_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(HERE, '..', 'include', 'viabel_b200.h'))
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src[:-3] + '.o')
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + ['-c', s, '-o', o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write('nvcc failed for %s:\n%s\n' % (src, out))
        else:
            with open(os.path.join(objdir, src[:-3] + '.ptxas.log'), 'w') as f:
                f.write(out)
            if verbose:
                print(out)
    if failed:
        raise RuntimeError('viabel_b200: CUDA build failed')
    if procs or force or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart', '-lcuda']
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
