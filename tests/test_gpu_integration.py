"""GPU: the integration oracle of SURVEY.md 8c(3) -- the reference's OWN lib/voxurf_fine.py (from baseline/_ref/lib, a
git-ignored copy placed by build()) executed unmodified twice: on the reference's own compiled CUDA kernels
(oracle/_ref) and on this repository's shim modules of the same names (voxurf_b200.render_utils_cuda,
total_variation_cuda, torch_scatter).  ret_dict and .grad must agree -- integers exactly -- and with the repository's
own model mirror too.  This is the drop-in claim of the C-ABI boundary, executed."""
import os

import numpy as np
import pytest
import torch

from voxurf_b200 import synthetic as S
from tests.helpers import T, product_fine_model

pytestmark = pytest.mark.gpu
RK = dict(near=0.3, far=6.0, bg=0.0, stepsize=0.5)


def _close(a, b, rtol=1e-5, atol=1e-6, msg=''):
    np.testing.assert_allclose(a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy(), rtol=rtol, atol=atol, err_msg=msg)


def _gclose(a, b, msg):
    b = b.detach().float().cpu()
    _close(a, b, 1e-4, 1e-4 * max(float(b.abs().max()), 1e-30), msg)


def _lin(seq):
    return [m for m in seq.modules() if isinstance(m, torch.nn.Linear)]


def test_reference_model_runs_unmodified_on_the_shim(tmp_path):
    from oracle import ref_model as RM
    if not RM.available():
        pytest.skip('baseline/_ref/lib not present (placed by __graft_entry__.build() in the build container)')
    from oracle import build_ref
    if build_ref.load_ref('render_utils_cuda') is None:
        pytest.skip('oracle/_ref not built')
    G, C, W, n_rays, step = 40, 6, 64, 700, 15003
    sc = S.make_fine_scene(G, C, W, seed=17)
    ckpt = os.path.join(tmp_path, 'mask.tar')
    RM.write_mask_ckpt(ckpt, sc['mask_density'], sc['mask_act_shift'], sc['mask_voxel_size_ratio'])
    cfg_model = {k: v for k, v in S.FINE_CFG.items() if k != 'stepsize'}
    ro, rd, vd = (T(x).cuda() for x in S.make_rays(n_rays, seed=3))
    target = T(S.make_target(vd.cpu().numpy())).cuda()
    out = {}
    for backend in ('ref', 'b200'):
        ns = RM.load_lib(backend)
        m = RM.build_fine(ns, G, C, W, ckpt, cfg_model, sc['sdf'], sc['k0'], (sc['rgbnet'], sc['k_rgbnet']))
        with RM.cuda_default():
            ret = m(ro, rd, vd, global_step=step, render_grad=True, render_depth=True, **RK)
            loss = torch.nn.functional.mse_loss(ret['rgb_marched'], target) + 0.5 * torch.nn.functional.mse_loss(ret['rgb_marched0'], target)
            loss = loss + 0.01 * m.density_total_variation(sdf_tv=0, smooth_grad_tv=0.05)
            loss.backward()
            m.sdf_total_variation_add_grad(0.01 * 0.1 / n_rays, True)
        out[backend] = (ret, loss.detach(), m.sdf.grid.grad.clone(), m.k0.grid.grad.clone(),
                        [l.weight.grad.clone() for l in _lin(m.rgbnet) + _lin(m.k_rgbnet)], m.nonempty_mask.clone())
    ra, rb = out['ref'][0], out['b200'][0]
    for k in ('mask', 'mask_outbbox'):
        assert torch.equal(ra[k], rb[k]), k
    assert torch.equal(out['ref'][5], out['b200'][5])
    for k in ('alphainv_cum', 'weights', 'rgb_marched', 'rgb_marched0', 'normal_marched', 'raw_alpha', 'raw_rgb', 'depth', 'gradient'):
        _close(rb[k], ra[k], 1e-5, 3e-6, k)
    _close(out['b200'][1], out['ref'][1], 1e-5, 1e-7, 'loss')
    _gclose(out['b200'][2], out['ref'][2], 'sdf grad'); _gclose(out['b200'][3], out['ref'][3], 'k0 grad')
    for a, b in zip(out['b200'][4], out['ref'][4]):
        _gclose(a, b, 'mlp weight grad')
    # and the repository's own model mirror (same constructor keywords, same ret_dict)
    mm = product_fine_model(sc)
    ret = mm(ro, rd, vd, global_step=step, render_grad=True, render_depth=True, **RK)
    for k in ('mask', 'mask_outbbox'):
        assert torch.equal(ra[k], ret[k]), k
    # (the mirror is pinned against the reference's CPU outputs at rtol 1e-5 in tests/test_gpu_model.py; here the other side
    # is ATen's CUDA grid_sample / sigmoid, and at 1/s ~ 320 the NeuS alpha amplifies last-bit sdf differences)
    errs = {k: float((ret[k].float() - ra[k].float()).abs().max()) for k in ('gradient', 'raw_alpha', 'alphainv_cum', 'weights', 'rgb_marched', 'rgb_marched0', 'normal_marched', 'depth')}
    print('mirror vs GPU reference, max abs err:', errs)
    for k in ('gradient', 'raw_alpha', 'alphainv_cum', 'weights', 'rgb_marched', 'rgb_marched0', 'normal_marched', 'depth'):
        _close(ret[k], ra[k], 5e-4, 2e-4, 'mirror ' + k)


def test_reference_womask_model_runs_unmodified_on_the_shim():
    """the unbounded ("womask") fine model, lib/voxurf_womask_fine.py:833-1071, reuses the same operators
    plus ub360_utils_cuda.cumdist_thres.  The reference's own file runs unmodified on the shim modules and gives the same
    ret_dict and gradients as on the reference's compiled kernels."""
    from oracle import ref_model as RM
    from oracle import build_ref
    if not RM.available() or build_ref.load_ref('render_utils_cuda') is None or build_ref.load_ref('ub360_utils_cuda') is None:
        pytest.skip('baseline/_ref/lib or oracle/_ref (incl. ub360_utils_cuda) not present')
    cfg = dict(num_voxels=32 ** 3, num_voxels_base=32 ** 3, num_voxels_bg=20 ** 3, alpha_init=1e-2, fast_color_thres=1e-4, bg_fast_color_thres=1e-4,
               posbase_pe=5, viewbase_pe=1, k_posbase_pe=5, k_viewbase_pe=1, k_res=True, rgbnet_depth=4, k_rgbnet_depth=4,
               k_grad_feat=(1.0,), k_sdf_feat=(), rgbnet_dim=6, rgbnet_width=32, center_sdf=True, k_center_sdf=False,
               grad_feat=(0.5, 1.0, 1.5, 2.0), sdf_feat=(0.5, 1.0, 1.5, 2.0), use_grad_norm=True, s_ratio=50, s_start=0.05,
               bg_rgbnet_dim=6, bg_rgbnet_width=32, bg_rgbnet_depth=3, smooth_ksize=0)
    n_rays = 256
    ro, rd, vd = (T(x).cuda() for x in S.make_rays(n_rays, seed=8, r_cam=2.0, r_target=0.4))
    target = T(S.make_target(vd.cpu().numpy())).cuda()
    out = {}
    for backend in ('ref', 'b200'):
        ns = RM.load_lib(backend, extra=('voxurf_womask_fine',))
        with RM.cuda_default():
            torch.manual_seed(5)
            m = ns.voxurf_womask_fine.Voxurf(xyz_min=torch.tensor([-1., -1., -1.]).cuda(), xyz_max=torch.tensor([1., 1., 1.]).cuda(), **cfg).cuda()
            with torch.no_grad():
                gen = torch.Generator(device='cuda').manual_seed(9)
                m.k0.grid.data.copy_(0.1 * torch.randn(m.k0.grid.shape, generator=gen, device='cuda'))
                m.bg_density.grid.data.copy_(0.5 * torch.randn(m.bg_density.grid.shape, generator=gen, device='cuda'))
            ret = m(ro, rd, vd, global_step=5003, near=0.2, far=4.0, bg=0.0, stepsize=0.5, inverse_y=False, flip_x=False, flip_y=False)
            loss = torch.nn.functional.mse_loss(ret['rgb_marched'], target)
            loss.backward()
        out[backend] = ({k: v.detach().clone() for k, v in ret.items() if torch.is_tensor(v)}, loss.detach(),
                        m.sdf.grid.grad.clone(), m.k0.grid.grad.clone(), m.bg_density.grid.grad.clone())
    ra, rb = out['ref'][0], out['b200'][0]
    assert set(ra) == set(rb) and 'rgb_marched' in ra
    for k in ra:
        if ra[k].dtype in (torch.bool, torch.int64, torch.int32):
            assert torch.equal(ra[k], rb[k]), k
        else:
            _close(rb[k], ra[k], 1e-5, 3e-6, k)
    _close(out['b200'][1], out['ref'][1], 1e-5, 1e-7, 'loss')
    for i, name in ((2, 'sdf grad'), (3, 'k0 grad'), (4, 'bg_density grad')):
        _gclose(out['b200'][i], out['ref'][i], name)
