"""CPU: the C-ABI library loads, exports every symbol include/voxurf_b200.h declares, and the host
binding refuses anything that is not a contiguous CUDA tensor (no compute here: no GPU)."""
import ctypes
import os

import pytest
import torch

from voxurf_b200 import _lib


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 38
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in protos:
        assert hasattr(lib, name), name
    assert _lib.library().vx_abi_version() == 1
    # reference surface: every pybind name of render_utils.cpp:170-184, total_variation.cpp:29-32, adam_upd.cpp:79-86
    from voxurf_b200 import adam_upd_cuda, render_utils_cuda, total_variation_cuda
    for fn in ['infer_t_minmax', 'infer_n_samples', 'infer_ray_start_dir', 'sample_pts_on_rays', 'sample_ndc_pts_on_rays',
               'sample_bg_pts_on_rays', 'maskcache_lookup', 'raw2alpha', 'raw2alpha_backward', 'raw2alpha_nonuni',
               'raw2alpha_nonuni_backward', 'alpha2weight', 'alpha2weight_backward']:
        assert callable(getattr(render_utils_cuda, fn))
    for fn in ['total_variation_add_grad', 'total_variation_add_grad_new']:
        assert callable(getattr(total_variation_cuda, fn))
    for fn in ['adam_upd', 'masked_adam_upd', 'adam_upd_with_perlr']:
        assert callable(getattr(adam_upd_cuda, fn))


def test_header_has_no_torch_types():
    src = open(_lib.HEADER).read()
    assert 'torch' not in src.replace('torch_scatter', '').replace('torch.', '') and 'at::' not in src


def test_cpu_tensors_are_rejected_loudly():
    from voxurf_b200 import render_utils_cuda
    a = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        render_utils_cuda.infer_t_minmax(a, a, torch.zeros(3), torch.ones(3), 0.1, 1.0)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, fs in os.walk(os.path.join(root, 'voxurf_b200')):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f
