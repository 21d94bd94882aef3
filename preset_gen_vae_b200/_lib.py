"""ctypes binding of libpgv.so (C ABI in include/pgv.h).

There is deliberately no fallback: if the shared library is missing or fails to load, or a call returns non-zero,
this module raises.  Prototypes are generated from include/pgv.h itself, so a signature change in the header that
is not mirrored by the callers fails at call time with a ctypes ArgumentError rather than corrupting memory.
"""
import ctypes
import os
import re
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libpgv.so')
HEADER_PATH = os.path.join(os.path.dirname(_PKG), 'include', 'pgv.h')

_CTYPES = {
    'int': ctypes.c_int, 'float': ctypes.c_float, 'size_t': ctypes.c_size_t, 'void': None,
    'int64_t': ctypes.c_int64, 'uint64_t': ctypes.c_uint64, 'double': ctypes.c_double,
    'pgv_stream_t': ctypes.c_void_p,
}


class PgvError(RuntimeError):
    pass


def parse_header(path=HEADER_PATH):
    """Returns {name: (restype, [argtypes])} for every function declared in pgv.h."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'^\s*#.*$', '', src, flags=re.M).replace('PGV_API', '')
    protos = {}
    for m in re.finditer(r'([A-Za-z_][\w\s\*]*?)\b(pgv_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()

        def ctype(decl):
            decl = decl.replace('const', ' ').strip()
            if '*' in decl:
                return ctypes.c_char_p if decl.startswith('char') and name == 'pgv_last_error' else ctypes.c_void_p
            base = decl.split()[0]
            return _CTYPES[base]
        argtypes = [] if args in ('', 'void') else [ctype(a) for a in args.split(',')]
        protos[name] = (ctype(ret + ' x') if '*' in ret else _CTYPES[ret.replace('const', '').strip()], argtypes)
    return protos


_lock = threading.Lock()
_lib = None
_handles = {}


def lib():
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PgvError("libpgv.so not found at %s: build it with `python -m preset_gen_vae_b200.csrc.build` "
                               "(there is no CPU / eager fallback)" % LIB_PATH)
            cdll = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes) in parse_header().items():
                fn = getattr(cdll, name)      # AttributeError if the .so does not export a declared symbol
                fn.restype = restype
                fn.argtypes = argtypes
            if cdll.pgv_version() != _header_version():
                raise PgvError("libpgv.so version %d does not match include/pgv.h (%d): rebuild" %
                               (cdll.pgv_version(), _header_version()))
            if os.environ.get('PGV_PDL') == '1':          # A/B switch (tools / bench): flow kernels WITH programmatic dependent launch
                cdll.pgv_debug_set_pdl(1)
            _lib = cdll
    return _lib


def _header_version():
    return int(re.search(r'#define\s+PGV_VERSION\s+(\d+)', open(HEADER_PATH).read()).group(1))


def check(rc, what=''):
    if rc != 0:
        raise PgvError("%s failed (code %d): %s" % (what or 'pgv call', rc, lib().pgv_last_error().decode()))


def handle(device=None):
    """Per-device pgv_handle (created on first use)."""
    if not torch.cuda.is_available():
        raise PgvError("no CUDA device: this library only runs on a B200 (sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    with _lock:
        h = _handles.get(idx)
    if h is None:
        h = ctypes.c_void_p()
        with torch.cuda.device(idx):
            check(lib().pgv_init(ctypes.byref(h), idx), 'pgv_init')
        with _lock:
            _handles[idx] = h
    return h


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device (or host) address of a contiguous tensor, or NULL for None."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_contiguous(), "pgv kernels take contiguous tensors"
    return ctypes.c_void_p(t.data_ptr())
