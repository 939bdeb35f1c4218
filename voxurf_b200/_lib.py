"""ctypes binding of libvoxurf_b200.so.

The argument types of every entry point are parsed from include/voxurf_b200.h, so the header is the
single source of truth for the C ABI.  `call(name, *args)` converts torch tensors to raw device
pointers (checking device, dtype and contiguity the way the reference's CHECK_INPUT does,
lib/cuda/render_utils.cpp:46-48), appends the current CUDA stream and raises RuntimeError on a
non-zero return code.  There is no fallback: a missing library is an error.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'voxurf_b200.h')
SO_PATH = os.environ.get('VX_SO') or os.path.join(_HERE, 'libvoxurf_b200.so')   # VX_SO: development builds (e.g. -DMC_TRACE)

_PTR_DTYPES = {
    'float': (torch.float32,), 'int': (torch.int32,), 'int64_t': (torch.int64,), 'bool': (torch.bool, torch.uint8),
    'uint8_t': (torch.uint8, torch.bool), 'uint32_t': (torch.int32,),
}
_SCALARS = {'float': ctypes.c_float, 'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'uint64_t': ctypes.c_uint64}


def parse_header(path=HEADER):
    """-> {name: (restype, [(kind, ctype_name, argname), ...])}; kind in {'ptr','host','scalar','stream'}"""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(int|unsigned long long|const char\*)\s+(vx_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if a.startswith('cudaStream_t'):
                    parsed.append(('stream', 'cudaStream_t', a.split()[-1]))
                    continue
                mm = re.match(r'(const )?(\w+)\s*(\*)?\s*(\w+)$', a)
                assert mm, (name, a)
                ctype, is_ptr, argname = mm.group(2), mm.group(3), mm.group(4)
                if is_ptr:
                    parsed.append(('host' if argname.endswith('_host') else 'ptr', ctype, argname))
                else:
                    parsed.append(('scalar', ctype, argname))
        protos[name] = (ret, parsed)
    return protos


_lib = None
_protos = None


def _load():
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f'{SO_PATH} is missing: build it with `python -m voxurf_b200.build` (or __graft_entry__.build()); '
            'voxurf_b200 has no CPU or PyTorch fallback')
    lib = ctypes.CDLL(SO_PATH)
    protos = parse_header()
    for name, (ret, args) in protos.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = {'int': ctypes.c_int, 'unsigned long long': ctypes.c_ulonglong, 'const char*': ctypes.c_char_p}[ret]
        fn.argtypes = [ctypes.c_void_p if k in ('ptr', 'host', 'stream') else _SCALARS[t] for k, t, _ in args]
    _lib, _protos = lib, protos
    return lib


def library():
    return _load()


def prototypes():
    _load()
    return _protos


def last_error():
    return _load().vx_last_error().decode()


def launch_count():
    return int(_load().vx_launch_count())


TIMING = None   # bench.py / profiling: a list collecting (name, start_event, end_event) of every call on its stream


def call(name, *args):
    if TIMING is not None and not torch.cuda.is_current_stream_capturing():
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        rc = _call(name, *args)
        ev[1].record()
        TIMING.append((name, ev[0], ev[1]))
        return rc
    return _call(name, *args)


def _call(name, *args):
    lib = _load()
    ret, spec = _protos[name]
    fn = getattr(lib, name)
    n_user = len([s for s in spec if s[0] != 'stream'])
    if len(args) != n_user:
        raise TypeError(f'{name} expects {n_user} arguments, got {len(args)}')
    cargs, keep = [], []
    it = iter(args)
    for kind, ctype, argname in spec:
        if kind == 'stream':
            cargs.append(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            continue
        a = next(it)
        if kind == 'ptr':
            if a is None:
                cargs.append(ctypes.c_void_p(0))
                continue
            if not torch.is_tensor(a):
                raise TypeError(f'{name}: {argname} must be a tensor or None')
            if not a.is_cuda:
                raise RuntimeError(f'{argname} must be a CUDA tensor')
            if not a.is_contiguous():
                raise RuntimeError(f'{argname} must be contiguous')
            if a.dtype not in _PTR_DTYPES[ctype]:
                raise RuntimeError(f'{name}: {argname} must have dtype {_PTR_DTYPES[ctype][0]}, got {a.dtype}')
            cargs.append(ctypes.c_void_p(a.data_ptr()))
        elif kind == 'host':
            if a is None:
                cargs.append(ctypes.c_void_p(0))
                continue
            seq = a.flatten().tolist() if torch.is_tensor(a) else a
            if ctype == 'float':
                buf = (ctypes.c_float * max(1, len(seq)))(*[float(v) for v in seq])
            else:
                buf = (_SCALARS[ctype] * max(1, len(seq)))(*[int(v) for v in seq])
            keep.append(buf)
            cargs.append(ctypes.cast(buf, ctypes.c_void_p))
        else:
            cargs.append(_SCALARS[ctype](a))
    rc = fn(*cargs)
    if ret == 'int' and spec and spec[-1][0] == 'stream' and rc != 0:
        raise RuntimeError(f'voxurf_b200 {name} failed (code {rc}): {last_error()}')
    return rc
