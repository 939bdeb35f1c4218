"""TEST INFRASTRUCTURE ONLY.  Builds the reference's own CUDA extensions into oracle/_ref/.

The reference sources are compiled *where they lie* under /root/reference/lib/cuda (never copied);
`oracle/ref_compat.h` is force-included to repair the one torch-2.x incompatibility.
Outputs (ninja files, objects, .so) go only to oracle/_ref/<module>/, which is git-ignored but
travels to the GPU box with the gpurun snapshot.  The resulting pybind modules
(render_utils_cuda, total_variation_cuda, adam_upd_cuda) are the GPU-side authority for the
bit-exact integer/index outputs (tests/test_gpu_vs_reference.py); they are never used by the
product path.

Usage: python oracle/build_ref.py [module ...]
"""
import os
import sys

REF = '/root/reference/lib/cuda'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
MODULES = {
    'render_utils_cuda': ['render_utils.cpp', 'render_utils_kernel.cu'],
    'total_variation_cuda': ['total_variation.cpp', 'total_variation_kernel.cu'],
    'adam_upd_cuda': ['adam_upd.cpp', 'adam_upd_kernel.cu'],
    'ub360_utils_cuda': ['ub360_utils.cpp', 'ub360_utils_kernel.cu'],
}


def so_path(name):
    return os.path.join(OUT, name, name + '.so')


def build(names=None, verbose=False):
    if not os.path.isdir(REF):
        return False
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    os.environ.setdefault('MAX_JOBS', '4')
    from torch.utils.cpp_extension import load
    shim = os.path.join(HERE, 'ref_compat.h')
    for name in (names or MODULES):
        if os.path.exists(so_path(name)):
            continue
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name,
             sources=[os.path.join(REF, s) for s in MODULES[name]],
             extra_cflags=['-include', shim, '-O2'],
             extra_cuda_cflags=['-include', shim, '-gencode', 'arch=compute_100a,code=sm_100a'],
             build_directory=bdir, verbose=verbose, is_python_module=False)
    return True


def load_ref(name):
    """Import a prebuilt reference module from oracle/_ref (no compilation). Returns None if absent."""
    p = so_path(name)
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (the .so links against libtorch)
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    ok = build(sys.argv[1:] or None, verbose=True)
    print('built' if ok else 'reference sources not present; nothing built')
