"""Build libmimo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mimo_b200.build [--force]

The library has no torch or Python dependency: it is a plain CUDA-runtime shared
object with the C-ABI of include/mimo_b200.h, loaded by mimo_b200/_lib.py via ctypes.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(ROOT, 'build')
LIB = os.path.join(HERE, 'libmimo_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _newest_dep():
    t = 0.0
    for d in (CSRC, os.path.join(ROOT, 'include')):
        for f in os.listdir(d):
            t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    dep = _newest_dep()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= dep:
        return LIB
    objs = []

    def compile_one(src):
        obj = os.path.join(BUILD, src[:-3] + '.o')
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcuda']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
