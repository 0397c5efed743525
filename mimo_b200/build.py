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


STAMP = LIB + '.stamp'


def _fingerprint():
    """sha256 over the flags and the CONTENT of every source / header the library is built from: the library is reused
    only when it was built from exactly this tree (file times mean nothing after a checkout or a snapshot copy)."""
    import hashlib
    h = hashlib.sha256(' '.join(FLAGS).encode())
    for d in (CSRC, os.path.join(ROOT, 'include')):
        for f in sorted(os.listdir(d)):
            h.update(f.encode())
            with open(os.path.join(d, f), 'rb') as fh:
                h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().split()[:1] == [fp]:
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
    ver = subprocess.run([NVCC, '--version'], capture_output=True, text=True).stdout.strip().splitlines()[-1:]
    with open(STAMP, 'w') as fh:
        fh.write('%s\nsources: %s\nflags: %s\nnvcc: %s\n' % (fp, ' '.join(sources()), ' '.join(FLAGS), ' '.join(ver)))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
