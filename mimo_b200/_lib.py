"""ctypes binding of libmimo_b200.so (the C-ABI of include/mimo_b200.h).

This is the stub INTEGRATION.md describes: plain pointers and sizes in, status
codes out.  Status codes are mapped to the exception classes the reference
raises on the same conditions: AssertionError for argument
/ label-range assertions, numpy.linalg.LinAlgError for a failed Cholesky.

There is NO fallback: if the shared object is missing or no sm_100 device is
present, every compute entry point raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmimo_b200.so')

OK, EINVAL, ECUDA, ENOTPD, EUNSUPPORTED = 0, 1, 2, 3, 4
F32, F64 = 0, 1
WRITE_RESP, WRITE_LSE, DRAW_LABELS, ACC_LSE = 1, 2, 4, 8

_T = {'i': ctypes.c_int, 'l': ctypes.c_int64, 'p': ctypes.c_void_p, 'z': ctypes.c_size_t, 'u': ctypes.c_uint64,
      'd': ctypes.c_double}

def _parse_header():
    """Read the prototypes from include/mimo_b200.h so the binding cannot drift from it:
    name -> (restype code, argument codes)."""
    import re
    path = os.path.join(os.path.dirname(HERE), 'include', 'mimo_b200.h')
    if not os.path.exists(path):
        raise ImportError('mimo_b200: the C-ABI header %s is missing -- the ctypes signatures are read from it; keep '
                          'include/ next to the package (in-tree layout) or copy mimo_b200.h there' % path)
    with open(path) as fh:
        text = fh.read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    sigs = {}
    for m in re.finditer(r'(const\s+char\s*\*|int|size_t)\s+(mimo_\w+)\s*\(([^)]*)\)\s*;', text):
        res, name, params = m.group(1), m.group(2), m.group(3).strip()
        codes = ''
        if params and params != 'void':
            for prm in params.split(','):
                prm = prm.strip()
                if '*' in prm:
                    codes += 'p'
                elif prm.startswith('int64_t'):
                    codes += 'l'
                elif prm.startswith('uint64_t'):
                    codes += 'u'
                elif prm.startswith('size_t'):
                    codes += 'z'
                elif prm.startswith('int'):
                    codes += 'i'
                elif prm.startswith('double'):
                    codes += 'd'
                else:
                    raise ValueError('unhandled parameter %r in %s' % (prm, name))
        sigs[name] = ('s' if '*' in res else ('z' if res == 'size_t' else 'i'), codes)
    return sigs


SIGNATURES = _parse_header()

_lib = None


class MimoCudaError(RuntimeError):
    pass


def load():
    """Load the shared object (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('mimo_b200: %s is missing -- run `python -m mimo_b200.build` '
                          '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_char_p if res == 's' else _T[res]
        fn.argtypes = [_T[a] for a in args]
    _lib = lib
    return lib


def exported_symbols():
    return list(SIGNATURES)


def last_error():
    return load().mimo_last_error_string().decode()


def check(status):
    if status == OK:
        return
    msg = last_error()
    if status == ENOTPD:
        raise np.linalg.LinAlgError(msg or 'Matrix is not positive definite')
    if status == EINVAL:
        raise AssertionError(msg)
    if status == EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise MimoCudaError(msg)


def call(name, *args):
    check(getattr(load(), name)(*args))


_device_checked = False


def require_device():
    """Fail loudly unless the extension is built and an sm_100 GPU is visible.  The positive
    answer is cached: cudaGetDeviceProperties costs milliseconds per call."""
    global _device_checked
    if _device_checked:
        return
    lib = load()
    if not lib.mimo_device_ok():
        raise MimoCudaError('mimo_b200 needs a Blackwell (sm_100) GPU: ' + last_error())
    _device_checked = True
